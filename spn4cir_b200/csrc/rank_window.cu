// Rank-window selection: rows [lo, hi) of  argsort(scores, descending)  for every row of a score block, without sorting
// the row.
//
// Reference: zscir/get_cir_data.py:33-34 `i2i_ranks = torch.argsort(sims_intra_i2i, descending=True)` followed by the
// slices `i2i_ranks[i][rank:rank_max]` of get_diff_id (:58-68; CIRR twin :106-109,131-141) -- README windows are
// [10000:20000) and [10000:15000) of 17k..47k images: a full N x N argsort of which only a band is used.
//
// One CTA per row.  Total order: score descending, ties -> lower column index (the rule of every ranking kernel here).
//   1. two radix selects (4 passes of 8 bits over the order-preserving 32-bit key) find the keys of rank lo and rank hi-1;
//   2. every column whose key lies between them is compacted into shared memory as (key << 32 | column) -- the window plus
//      the ties of its two boundary values;
//   3. one bitonic sort of that band; ranks [lo, hi) are copied out.
// HBM/L2-bound streaming of the row (10 passes, the block of rows is L2-resident) + an in-smem sort of <= 16384 entries.
#include <cuda_runtime.h>
#include <stdint.h>

#include "sim_rank.cuh"

namespace spn {
namespace {

constexpr int kRwThreads = 1024;
constexpr int kRwCap = 16384;  // band capacity (entries of 8 bytes: 128 KiB of shared memory)

// ascending key order == descending float order (NaN sorts last)
__device__ __forceinline__ uint32_t desc_key(float x) {
  uint32_t b = __float_as_uint(x);
  if (x != x) return 0xFFFFFFFFu;
  if (x == 0.f) b = 0u;  // -0.0 == +0.0: equal scores are ordered by column, as a stable sort does
  const uint32_t u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // ascending-monotone
  return ~u;
}

// key of the element with 0-based rank `r` in ascending key order, and the number of keys strictly below it
__device__ uint32_t radix_select(const float* __restrict__ row, int G, int r, uint32_t* hist, uint32_t* bcast, int* below_out) {
  uint32_t prefix = 0, mask = 0;
  int below = 0;  // elements whose key is < the current prefix range
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int c = threadIdx.x; c < G; c += blockDim.x) {
      const uint32_t k = desc_key(row[c]);
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = below;
      uint32_t d = 0;
      for (; d < 256; ++d) {
        if (acc + int(hist[d]) > r) break;
        acc += int(hist[d]);
      }
      if (d > 255) d = 255;  // (r >= G: cannot happen, the host clamps)
      bcast[0] = d;
      bcast[1] = uint32_t(acc);
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    mask |= 255u << shift;
    below = int(bcast[1]);
    __syncthreads();
  }
  *below_out = below;
  return prefix;
}

__global__ void __launch_bounds__(kRwThreads, 1)
rank_window_kernel(const float* __restrict__ scores, int R, int G, long long pitch, int lo, int hi, int32_t* __restrict__ out,
                   int* __restrict__ overflow) {
  extern __shared__ unsigned long long band[];  // [cap]
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bcast[2];
  __shared__ int n_band;
  const int r = blockIdx.x;
  if (r >= R) return;
  const float* row = scores + size_t(r) * pitch;
  const int W = hi - lo;
  int below_lo = 0, below_hi = 0;
  const uint32_t k_lo = radix_select(row, G, lo, hist, bcast, &below_lo);
  const uint32_t k_hi = radix_select(row, G, hi - 1, hist, bcast, &below_hi);
  (void)below_hi;
  if (threadIdx.x == 0) n_band = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < G; c += blockDim.x) {
    const uint32_t k = desc_key(row[c]);
    if (k >= k_lo && k <= k_hi) {
      const int pos = atomicAdd(&n_band, 1);
      if (pos < kRwCap) band[pos] = (static_cast<unsigned long long>(k) << 32) | uint32_t(c);
    }
  }
  __syncthreads();
  const int n = n_band;
  if (n > kRwCap) {  // more ties on the window's boundary values than the band holds: the host takes the sort path
    if (threadIdx.x == 0) atomicExch(overflow, 1);
    return;
  }
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) band[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = band[i], b = band[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            band[i] = b;
            band[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  // band[0] has rank below_lo; the window starts at offset lo - below_lo
  const int off = lo - below_lo;
  for (int j = threadIdx.x; j < W; j += blockDim.x) out[size_t(r) * W + j] = int32_t(uint32_t(band[off + j] & 0xFFFFFFFFull));
}

}  // namespace

int rank_window(const float* scores, int R, int G, long long pitch, int lo, int hi, int32_t* out, int* overflow,
                cudaStream_t stream) {
  SPN_REQUIRE(R > 0 && G > 0 && pitch >= G, SPN_ERR_BAD_SHAPE, "rank_window: bad shape R=%d G=%d pitch=%lld", R, G, pitch);
  SPN_REQUIRE(lo >= 0 && hi > lo && hi <= G, SPN_ERR_BAD_ARG, "rank_window: need 0 <= lo < hi <= G (lo=%d hi=%d G=%d)", lo, hi, G);
  SPN_REQUIRE(hi - lo <= kRwCap * 3 / 4, SPN_ERR_BAD_ARG, "rank_window: window wider than %d ranks per call (split it)", kRwCap * 3 / 4);
  const size_t smem = size_t(kRwCap) * sizeof(unsigned long long);
  SPN_CUDA_CHECK(cudaFuncSetAttribute(rank_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  SPN_CUDA_CHECK(cudaMemsetAsync(overflow, 0, sizeof(int), stream));
  rank_window_kernel<<<R, kRwThreads, smem, stream>>>(scores, R, G, pitch, lo, hi, out, overflow);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
