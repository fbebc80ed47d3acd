// Warp-resident sorted top-K list (K <= 64) shared by the retrieval kernels (sim_rank.cu, sim_tc.cu).
// Ordering rule everywhere: descending score, ties -> lower global index first.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>

namespace spn {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ bool sorts_before(float av, int ai, float bv, int bi) {
  return (av > bv) || (av == bv && ai < bi);
}

// Insert (cv, ci) into the descending list held as positions {lane} (v0,i0) and {lane+32} (v1,i1).
__device__ __forceinline__ void topk_insert(float cv, int ci, float& v0, int& i0, float& v1, int& i1, int lane) {
  const int pos = __popc(__ballot_sync(kFull, sorts_before(v0, i0, cv, ci))) +
                  __popc(__ballot_sync(kFull, sorts_before(v1, i1, cv, ci)));
  float pv0 = __shfl_up_sync(kFull, v0, 1);
  int pi0 = __shfl_up_sync(kFull, i0, 1);
  float pv1 = __shfl_up_sync(kFull, v1, 1);
  int pi1 = __shfl_up_sync(kFull, i1, 1);
  const float v0_last = __shfl_sync(kFull, v0, 31);
  const int i0_last = __shfl_sync(kFull, i0, 31);
  if (lane == 0) {
    pv1 = v0_last;
    pi1 = i0_last;
  }
  if (lane == pos) {
    v0 = cv;
    i0 = ci;
  } else if (lane > pos) {
    v0 = pv0;
    i0 = pi0;
  }
  if (lane + 32 == pos) {
    v1 = cv;
    i1 = ci;
  } else if (lane + 32 > pos) {
    v1 = pv1;
    i1 = pi1;
  }
}

// Offer one candidate per lane (cv, ci, valid) to the K-long list.
__device__ __forceinline__ void topk_offer(float cv, int ci, bool valid, int K, float& v0, int& i0, float& v1,
                                           int& i1, int lane) {
  float lv = (K - 1 < 32) ? __shfl_sync(kFull, v0, K - 1) : __shfl_sync(kFull, v1, K - 33);
  int li = (K - 1 < 32) ? __shfl_sync(kFull, i0, K - 1) : __shfl_sync(kFull, i1, K - 33);
  unsigned mask = __ballot_sync(kFull, valid && sorts_before(cv, ci, lv, li));
  while (mask) {
    const int src = __ffs(mask) - 1;
    const float bv = __shfl_sync(kFull, cv, src);
    const int bi = __shfl_sync(kFull, ci, src);
    topk_insert(bv, bi, v0, i0, v1, i1, lane);
    lv = (K - 1 < 32) ? __shfl_sync(kFull, v0, K - 1) : __shfl_sync(kFull, v1, K - 33);
    li = (K - 1 < 32) ? __shfl_sync(kFull, i0, K - 1) : __shfl_sync(kFull, i1, K - 33);
    mask = __ballot_sync(kFull, valid && lane > src && sorts_before(cv, ci, lv, li));
  }
}

}  // namespace spn
