// Sampled-negatives contrastive loss: every query scores its OWN short candidate list
// [positive, negative_1 .. negative_k] gathered from the resident bank, CE with label 0, mean over the batch.
//
// Reference: CIRPlus.part_infonce_loss, clip4cir/models.py:115-128 (candidate ids from get_neg_id :109-113, the
// per-sample infonce_loss call :125-126); zscir/models_bank.py:103-122 is the same loop.  The candidate ids are drawn on
// the host with the reference's own RNG calls (spn4cir_b200/compat.py) -- this kernel takes them as an index matrix.
//
// Not a GEMM: B x (k + 1) independent D-long dot products against gathered rows -> HBM/L2-latency bound gather.  One
// CTA per query: warps take candidates round-robin (coalesced bf16x2 row reads, shuffle reduction), the logits live in
// shared memory, one block-wide soft-max, then every thread owns a few columns of dq and walks the candidates again.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "bank_ce.cuh"

namespace spn {
namespace {

constexpr int kPartThreads = 128;

template <typename SrcT>
__device__ __forceinline__ float part_load(const SrcT* p, size_t i);
template <>
__device__ __forceinline__ float part_load<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float part_load<__half>(const __half* p, size_t i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ float part_load<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  __syncthreads();  // red may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < kPartThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <typename SrcT>
__global__ void __launch_bounds__(kPartThreads)
part_ce_kernel(const SrcT* __restrict__ q_src, const __nv_bfloat16* __restrict__ bank, const long long* __restrict__ cand,
               int B, int K1, int D, long long N, float inv_tau, float gscale, float* __restrict__ loss_q,
               float* __restrict__ loss_mean, float* __restrict__ dq, float* __restrict__ dlogit_dot) {
  extern __shared__ float smem[];
  float* logit = smem;            // [K1]
  float* qs = smem + K1;          // [D] query row rounded to bf16 (the arithmetic type of the path)
  __shared__ float red[kPartThreads / 32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < D; d += kPartThreads)
    qs[d] = __bfloat162float(__float2bfloat16_rn(part_load<SrcT>(q_src, size_t(b) * D + d)));
  __syncthreads();
  // logits: one warp per candidate
  for (int c = warp; c < K1; c += kPartThreads / 32) {
    const long long r = cand[size_t(b) * K1 + c];
    float acc = 0.f;
    if (r >= 0 && r < N) {
      const __nv_bfloat16* row = bank + r * D;
      for (int d = lane * 2; d < D; d += 64) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + d));
        acc = fmaf(qs[d], v.x, fmaf(qs[d + 1], v.y, acc));
      }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logit[c] = (r >= 0 && r < N) ? acc * inv_tau : -INFINITY;  // out-of-range ids are masked out
  }
  __syncthreads();
  float m = -INFINITY;
  for (int c = threadIdx.x; c < K1; c += kPartThreads) m = fmaxf(m, logit[c]);
  m = block_reduce(m, true, red);
  float l = 0.f;
  for (int c = threadIdx.x; c < K1; c += kPartThreads) l += expf(logit[c] - m);
  l = block_reduce(l, false, red);
  const float lse = m + logf(l);
  const float s0 = logit[0];
  if (threadIdx.x == 0) {
    const float loss = lse - s0;  // CE with label 0 (models.py:125-126)
    if (loss_q) loss_q[b] = loss;
    atomicAdd(loss_mean, loss / float(B));
  }
  if (dlogit_dot) {  // sum_c dlogit_c s_c  (d loss / d tau = -(1 / (tau B)) sum_b of this)
    float e = 0.f;
    for (int c = threadIdx.x; c < K1; c += kPartThreads) {
      const float s = logit[c];
      if (s > -INFINITY) e += (expf(s - lse) - (c == 0 ? 1.f : 0.f)) * s;
    }
    e = block_reduce(e, false, red);
    if (threadIdx.x == 0) dlogit_dot[b] = e;
  }
  if (!dq) return;
  __syncthreads();
  // p_c - [c == 0] in place, then dq[b, d] = gscale * sum_c dlogit_c bank[cand_c, d]
  for (int c = threadIdx.x; c < K1; c += kPartThreads) {
    const float s = logit[c];
    logit[c] = (s > -INFINITY) ? expf(s - lse) - (c == 0 ? 1.f : 0.f) : 0.f;
  }
  __syncthreads();
  for (int d0 = threadIdx.x * 2; d0 < D; d0 += kPartThreads * 2) {
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < K1; ++c) {
      const float w = logit[c];
      if (w == 0.f) continue;
      const long long r = cand[size_t(b) * K1 + c];
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bank + r * D + d0));
      a0 = fmaf(w, v.x, a0);
      a1 = fmaf(w, v.y, a1);
    }
    *reinterpret_cast<float2*>(dq + size_t(b) * D + d0) = make_float2(a0 * gscale, a1 * gscale);
  }
}

}  // namespace

int part_ce(const void* q_src, int dtype, const void* bank, const long long* cand, float inv_tau, int B, int K1, int D,
            long long N, float* loss_q, float* loss_mean, float* dq, float* dlogit_dot, cudaStream_t stream) {
  SPN_REQUIRE(B > 0 && K1 >= 1 && N > 0, SPN_ERR_BAD_SHAPE, "part_ce: bad shape B=%d K1=%d N=%lld", B, K1, N);
  SPN_REQUIRE(D >= 2 && D % 2 == 0 && D <= 4096, SPN_ERR_BAD_SHAPE, "part_ce: D=%d must be even and <= 4096", D);
  SPN_REQUIRE(K1 <= 8192, SPN_ERR_BAD_SHAPE, "part_ce: at most 8192 candidates per query (got %d)", K1);
  SPN_REQUIRE(inv_tau > 0.f && isfinite(inv_tau), SPN_ERR_BAD_ARG, "inv_tau must be positive and finite (got %g)", inv_tau);
  SPN_CUDA_CHECK(cudaMemsetAsync(loss_mean, 0, sizeof(float), stream));
  const size_t smem = size_t(K1 + D) * sizeof(float);
  const __nv_bfloat16* bk = static_cast<const __nv_bfloat16*>(bank);
  const float gscale = inv_tau / float(B);
#define SPN_PART_LAUNCH(TYPE)                                                                                         \
  do {                                                                                                                \
    SPN_CUDA_CHECK(cudaFuncSetAttribute(part_ce_kernel<TYPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))); \
    part_ce_kernel<TYPE><<<B, kPartThreads, smem, stream>>>(static_cast<const TYPE*>(q_src), bk, cand, B, K1, D, N,   \
                                                           inv_tau, gscale, loss_q, loss_mean, dq, dlogit_dot);       \
  } while (0)
  switch (dtype) {
    case SPN_F32: SPN_PART_LAUNCH(float); break;
    case SPN_F16: SPN_PART_LAUNCH(__half); break;
    case SPN_BF16: SPN_PART_LAUNCH(__nv_bfloat16); break;
    default: return fail(SPN_ERR_BAD_DTYPE, "part_ce: unknown dtype %d", dtype);
  }
#undef SPN_PART_LAUNCH
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
