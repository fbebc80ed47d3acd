// Row-sharded bank: the small per-rank kernels either side of the exchange.
//
// Reference anchor: zscir/models_bank.py:90-101,130-134 + zscir/train_bank.py:98-120 (one loss over ALL bank rows); here
// every rank owns N / R rows and scores all R * B queries of the step against them with the fused sweep
// (bank_ce_tc.cu).  What travels is additive: o_acc [Bt, D], and per query (l, s_t, es, owned).
//
//   shard_pack : (l_acc, s_t, es_acc, lbl_row) -> vec [4, Bt] with the owner-only entries zeroed elsewhere, so ONE
//                sum all-reduce merges them (and `owned` counts the ranks that hold the label: 0 = label out of range).
//   shard_fold : after that all-reduce the rank that owns label b folds the target-row terms of the gradient into its
//                partial:  o_acc[b] += (-l_b + eps L_b) bank[target row],  L_b = l_b + exp(s_t - ref)
//                so that sum_r o_r = o - l bank_t + eps L bank_t and the finalize step needs no bank_t: the second
//                [Bt, D] reduce-scatter of round 1 (a mostly-zero fp32 matrix) is gone.
#include <cuda_bf16.h>
#include <math.h>

#include "bank_ce.cuh"

namespace spn {
namespace {

__global__ void shard_pack_kernel(const float* __restrict__ l_acc, const float* __restrict__ s_t,
                                  const float* __restrict__ es_acc, const long long* __restrict__ lbl_row, int Bt,
                                  float* __restrict__ vec) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Bt) return;
  const bool own = lbl_row[b] >= 0;
  vec[b] = l_acc[b];
  vec[size_t(Bt) + b] = own ? s_t[b] : 0.f;
  vec[size_t(2) * Bt + b] = es_acc ? es_acc[b] : 0.f;
  vec[size_t(3) * Bt + b] = own ? 1.f : 0.f;
}

// one warp per query
__global__ void shard_fold_kernel(float* __restrict__ o_acc, const float* __restrict__ vec, const float* __restrict__ ref,
                                  const long long* __restrict__ lbl_row, const __nv_bfloat16* __restrict__ bank, int Bt,
                                  int D, float eps, float* __restrict__ s_t_out) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= Bt) return;
  const float l = vec[b], st = vec[size_t(Bt) + b], owned = vec[size_t(3) * Bt + b];
  if (lane == 0) s_t_out[b] = owned > 0.5f ? st : -INFINITY;  // a label no rank owns: loss = +inf, as on one GPU
  const long long r = lbl_row[b];
  if (r < 0) return;
  const float L = l + expf(st - ref[b]);
  const float coef = -l + eps * L;
  const __nv_bfloat16* row = bank + size_t(r) * D;
  float* o = o_acc + size_t(b) * D;
  for (int d = lane * 2; d < D; d += 64) {
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + d));
    float2 cur = *reinterpret_cast<float2*>(o + d);
    cur.x = fmaf(coef, v.x, cur.x);
    cur.y = fmaf(coef, v.y, cur.y);
    *reinterpret_cast<float2*>(o + d) = cur;
  }
}

}  // namespace

int shard_pack(const float* l_acc, const float* s_t, const float* es_acc, const long long* lbl_row, int Bt, float* vec,
               cudaStream_t stream) {
  SPN_REQUIRE(Bt > 0, SPN_ERR_BAD_SHAPE, "shard_pack: Bt must be positive");
  shard_pack_kernel<<<(Bt + 255) / 256, 256, 0, stream>>>(l_acc, s_t, es_acc, lbl_row, Bt, vec);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int shard_fold(float* o_acc, const float* vec, const float* ref, const long long* lbl_row, const void* bank, int Bt,
               int D, float eps, float* s_t_out, cudaStream_t stream) {
  SPN_REQUIRE(Bt > 0 && D > 0 && D % 2 == 0, SPN_ERR_BAD_SHAPE, "shard_fold: bad shape Bt=%d D=%d", Bt, D);
  shard_fold_kernel<<<(Bt + 7) / 8, 256, 0, stream>>>(o_acc, vec, ref, lbl_row, static_cast<const __nv_bfloat16*>(bank), Bt,
                                                       D, eps, s_t_out);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
