// Prologue kernels: dtype cast to the bf16 operand format and the fused query assembly
// normalize(refer_bank[idx] + text)  (clip4cir/models_negplus.py:133-137).  Pure HBM-bound
// streaming kernels: 16-byte vector loads, grid sized as a multiple of the SM count.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "bank_ce.cuh"
#include "ptx.cuh"

namespace spn {
namespace {

template <typename SrcT>
__device__ __forceinline__ float to_f32(SrcT v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// 8 elements per thread per iteration: one or two 16-byte loads, one 16-byte store.
template <typename SrcT>
__global__ void cast_bf16_kernel(const SrcT* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n8 = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    if (sizeof(SrcT) == 4) {
      const float4 a = reinterpret_cast<const float4*>(src)[2 * i];
      const float4 b = reinterpret_cast<const float4*>(src)[2 * i + 1];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      const uint4 raw = reinterpret_cast<const uint4*>(src)[i];
      const SrcT* h = reinterpret_cast<const SrcT*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = to_f32<SrcT>(h[j]);
    }
    uint4 out;
    out.x = pack_bf16x2(v[0], v[1]);
    out.y = pack_bf16x2(v[2], v[3]);
    out.z = pack_bf16x2(v[4], v[5]);
    out.w = pack_bf16x2(v[6], v[7]);
    reinterpret_cast<uint4*>(dst)[i] = out;
  }
  // tail (< 8 elements)
  if (blockIdx.x == 0 && threadIdx.x < n - n8 * 8) {
    const long long i = n8 * 8 + threadIdx.x;
    dst[i] = __float2bfloat16_rn(to_f32<SrcT>(src[i]));
  }
}

// one warp per query row
__global__ void query_assemble_kernel(const float* __restrict__ refer_bank, const long long* __restrict__ refer_idx,
                                      const float* __restrict__ text, int B, int D, long long R,
                                      __nv_bfloat16* __restrict__ q_bf16, float* __restrict__ q_f32,
                                      float* __restrict__ inv_norm) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long ri = refer_idx[b];
  if (ri < 0 || ri >= R) {  // the reference raises an IndexError here: make it loud (NaN rows), never read out of bounds
    const float nan = __int_as_float(0x7fc00000);
    for (int d = lane; d < D; d += 32) {
      q_bf16[size_t(b) * D + d] = __float2bfloat16_rn(nan);
      if (q_f32) q_f32[size_t(b) * D + d] = nan;
    }
    if (lane == 0 && inv_norm) inv_norm[b] = nan;
    return;
  }
  const float* r = refer_bank + size_t(ri) * D;
  const float* t = text + size_t(b) * D;
  float ss = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(r + d);
    const float4 c = *reinterpret_cast<const float4*>(t + d);
    const float x0 = a.x + c.x, x1 = a.y + c.y, x2 = a.z + c.z, x3 = a.w + c.w;
    ss += x0 * x0 + x1 * x1 + x2 * x2 + x3 * x3;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
  if (lane == 0 && inv_norm) inv_norm[b] = inv;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(r + d);
    const float4 c = *reinterpret_cast<const float4*>(t + d);
    const float x0 = (a.x + c.x) * inv, x1 = (a.y + c.y) * inv, x2 = (a.z + c.z) * inv, x3 = (a.w + c.w) * inv;
    uint2 o2;
    o2.x = pack_bf16x2(x0, x1);
    o2.y = pack_bf16x2(x2, x3);
    *reinterpret_cast<uint2*>(q_bf16 + size_t(b) * D + d) = o2;
    if (q_f32) *reinterpret_cast<float4*>(q_f32 + size_t(b) * D + d) = make_float4(x0, x1, x2, x3);
  }
}

// one warp per query row: backward of q = x / |x| (x = refer_bank[idx] + text), i.e. the gradient w.r.t. text:
//   dtext = g * inv_norm * (dq - q <q, dq>)
__global__ void query_assemble_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ q_f32,
                                          const float* __restrict__ inv_norm, const float* __restrict__ gscale_dev, int B,
                                          int D, float* __restrict__ dtext) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* g = dq + size_t(b) * D;
  const float* q = q_f32 + size_t(b) * D;
  float dot = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(g + d);
    const float4 c = *reinterpret_cast<const float4*>(q + d);
    dot += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
  }
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float s = inv_norm[b] * (gscale_dev ? __ldg(gscale_dev) : 1.f);
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(g + d);
    const float4 c = *reinterpret_cast<const float4*>(q + d);
    *reinterpret_cast<float4*>(dtext + size_t(b) * D + d) =
        make_float4(s * (a.x - c.x * dot), s * (a.y - c.y * dot), s * (a.z - c.z * dot), s * (a.w - c.w * dot));
  }
}

}  // namespace

int cast_to_bf16(const void* src, int src_dtype, void* dst, long long n, cudaStream_t stream) {
  SPN_REQUIRE(n >= 0, SPN_ERR_BAD_SHAPE, "cast_to_bf16: negative element count");
  if (n == 0) return SPN_OK;
  SPN_REQUIRE(aligned16(src) && aligned16(dst), SPN_ERR_BAD_ALIGN, "cast_to_bf16: pointers must be 16-byte aligned");
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)sms * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  __nv_bfloat16* d = static_cast<__nv_bfloat16*>(dst);
  switch (src_dtype) {
    case SPN_F32: cast_bf16_kernel<float><<<int(blocks), 256, 0, stream>>>(static_cast<const float*>(src), d, n); break;
    case SPN_F16: cast_bf16_kernel<__half><<<int(blocks), 256, 0, stream>>>(static_cast<const __half*>(src), d, n); break;
    case SPN_BF16:
      SPN_CUDA_CHECK(cudaMemcpyAsync(dst, src, size_t(n) * 2, cudaMemcpyDeviceToDevice, stream));
      return SPN_OK;
    default: return fail(SPN_ERR_BAD_DTYPE, "cast_to_bf16: unknown source dtype %d", src_dtype);
  }
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int query_assemble(const float* refer_bank, const long long* refer_idx, const float* text, int B, int D, long long R,
                   void* q_bf16, float* q_f32, float* inv_norm, cudaStream_t stream) {
  SPN_REQUIRE(B > 0 && D > 0 && D % 4 == 0, SPN_ERR_BAD_SHAPE, "query_assemble: need B > 0 and D %% 4 == 0 (B=%d D=%d)", B, D);
  SPN_REQUIRE(aligned16(refer_bank) && aligned16(text) && aligned16(q_bf16), SPN_ERR_BAD_ALIGN,
              "query_assemble: pointers must be 16-byte aligned");
  const int warps_per_block = 8;
  query_assemble_kernel<<<(B + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      refer_bank, refer_idx, text, B, D, R, static_cast<__nv_bfloat16*>(q_bf16), q_f32, inv_norm);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int query_assemble_bwd(const float* dq, const float* q_f32, const float* inv_norm, const float* gscale_dev, int B, int D,
                       float* dtext, cudaStream_t stream) {
  SPN_REQUIRE(B > 0 && D > 0 && D % 4 == 0, SPN_ERR_BAD_SHAPE, "query_assemble_bwd: need B > 0 and D %% 4 == 0 (B=%d D=%d)", B, D);
  SPN_REQUIRE(aligned16(dq) && aligned16(q_f32) && aligned16(dtext), SPN_ERR_BAD_ALIGN,
              "query_assemble_bwd: pointers must be 16-byte aligned");
  query_assemble_bwd_kernel<<<(B + 7) / 8, 256, 0, stream>>>(dq, q_f32, inv_norm, gscale_dev, B, D, dtext);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
