// One-CTA tcgen05 GEMM used by tests/test_umma_selftest.py to pin, on hardware, every encoding the
// bank-CE kernels rely on: TMA 128B-swizzled boxes, K-major and MN-major shared-memory descriptors,
// the bf16 instruction descriptor, A-operand-from-TMEM packing, and the 32x32b TMEM load/store map.
//   C[128, N] = A[128, K] * B^T   (B K-major  [N, K])      mode bit1 = 0
//   C[128, N] = A[128, K] * B     (B MN-major [K, N])      mode bit1 = 1 (K == 128)
//   mode bit0 = 1: A is first written to TMEM as packed bf16 (tcgen05.st) and read by the MMA from there.
#include <cuda_bf16.h>

#include "bank_ce.cuh"
#include "ptx.cuh"

namespace spn {
namespace {

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const uint32_t* __restrict__ a_gmem_u32, float* __restrict__ c, int N, int K, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~uintptr_t(1023u);
  uint8_t* a_smem = reinterpret_cast<uint8_t*>(base);       // up to 4 blocks of 16 KiB
  uint8_t* b_smem = a_smem + 4 * kBlockBytes;               // up to 4 blocks of 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + 4 * kBlockBytes);
  uint64_t* ld_full = bars;
  uint64_t* mma_done = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const bool a_tmem = mode & 1, b_mn = mode & 2;
  const int kblocks = K / 64;

  if (threadIdx.x == 0) {
    mbar_init(ld_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_row = tmem_base + ((warp * 32u) << 16);
  const int row = int(warp * 32u + lane);

  if (a_tmem) {
    // thread == row: pack its K bf16 (K/2 words) into TMEM columns [256, 256 + K/2)
    uint32_t u[32];
    for (int c = 0; c < K / 64; ++c) {
#pragma unroll
      for (int j = 0; j < 32; ++j) u[j] = a_gmem_u32[size_t(row) * (K / 2) + c * 32 + j];
      tmem_st_32x32b_x32(t_row + 256u + uint32_t(c) * 32u, u);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    const int b_blocks = b_mn ? N / 64 : kblocks;
    const uint32_t b_box_rows = b_mn ? 128u : uint32_t(N);
    const uint32_t bytes = uint32_t(kblocks) * kBlockBytes + uint32_t(b_blocks) * b_box_rows * 128u;
    mbar_arrive_expect_tx(ld_full, bytes);
    for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(a_smem + size_t(kb) * kBlockBytes, &tmap_a, ld_full, kb * 64, 0);
    for (int j = 0; j < b_blocks; ++j) tma_load_2d(b_smem + size_t(j) * kBlockBytes, &tmap_b, ld_full, j * 64, 0);
    mbar_wait(ld_full, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(a_smem), b_addr = smem_u32(b_smem);
    if (!b_mn) {
      const uint32_t idesc = make_idesc_bf16_f32(128, N, 0, 0);
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint64_t adesc = make_smem_desc_sw128(a_addr + uint32_t(kb) * kBlockBytes, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(b_addr + uint32_t(kb) * kBlockBytes, 16, 1024);
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (kb | k) != 0;
          if (a_tmem)
            umma_ts_f16(tmem_base, tmem_base + 256u + uint32_t(kb * 4 + k) * 8u, bdesc + uint64_t(k * 2), idesc, acc);
          else
            umma_ss_f16(tmem_base, adesc + uint64_t(k * 2), bdesc + uint64_t(k * 2), idesc, acc);
        }
      }
    } else if (mode & 4) {
      // one wide MN-major instruction: N = 64 * b_blocks columns, the 64-column atoms one block (LBO) apart
      const uint32_t idesc = make_idesc_bf16_f32(128, N, 0, 1);
      const uint64_t bdesc = make_smem_desc_sw128(b_addr, kBlockBytes, 1024);
      for (int k = 0; k < 8; ++k) {
        const uint64_t adesc = make_smem_desc_sw128(a_addr + uint32_t(k / 4) * kBlockBytes, 16, 1024) + uint64_t((k % 4) * 2);
        if (a_tmem)
          umma_ts_f16(tmem_base, tmem_base + 256u + uint32_t(k) * 8u, bdesc + uint64_t(k * 128), idesc, k != 0);
        else
          umma_ss_f16(tmem_base, adesc, bdesc + uint64_t(k * 128), idesc, k != 0);
      }
    } else {
      const uint32_t idesc = make_idesc_bf16_f32(128, 64, 0, 1);
      for (int j = 0; j < N / 64; ++j) {
        const uint64_t bdesc = make_smem_desc_sw128(b_addr + uint32_t(j) * kBlockBytes, kBlockBytes, 1024);
        for (int k = 0; k < 8; ++k) {  // K = 128 = 8 steps of 16 rows (2048 B each)
          const uint64_t adesc = make_smem_desc_sw128(a_addr + uint32_t(k / 4) * kBlockBytes, 16, 1024) + uint64_t((k % 4) * 2);
          if (a_tmem)
            umma_ts_f16(tmem_base + uint32_t(j) * 64u, tmem_base + 256u + uint32_t(k) * 8u, bdesc + uint64_t(k * 128), idesc, k != 0);
          else
            umma_ss_f16(tmem_base + uint32_t(j) * 64u, adesc, bdesc + uint64_t(k * 128), idesc, k != 0);
        }
      }
    }
    umma_commit(mma_done);
  }
  __syncwarp();
  mbar_wait(mma_done, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int cb = 0; cb < N / 32; ++cb) {
    tmem_ld_32x32b_x32(t_row + uint32_t(cb) * 32u, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) c[size_t(row) * N + cb * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- raw MMA issue-rate probe: every CTA issues `iters` x 4 k-steps of a PATTERN of up to 8 MMAs back to back from
// resident smem tiles (contents irrelevant).  One pattern op per byte of `pattern` (0 terminates):
//   bit0 valid | bit1 A from TMEM | bits2-3 N (0: 64, 1: 128, 2: 256) | bit4 B MN-major | bits5-6 B region (16 KiB
//   each) | bit7 second accumulator.  `group` > 1 repeats each op `group` times (k-steps) before moving on, i.e. it
// controls how often the instruction kind switches.  Reports cycles per CTA.
__global__ void __launch_bounds__(128, 1)
umma_bench_kernel(unsigned long long pattern, int iters, int group, unsigned long long* __restrict__ out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5;
  uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~uintptr_t(1023u);
  uint8_t* a_smem = reinterpret_cast<uint8_t*>(base);                // 16 KiB
  uint8_t* b_smem = a_smem + kBlockBytes;                            // 4 regions x 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + 4 * kBlockBytes);
  uint64_t* done = bars;        // MMA completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  for (uint32_t i = threadIdx.x; i < 5 * kBlockBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_smem)[i] = 0;
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(a_smem), 16, 1024);
    const uint64_t bdesc_k = make_smem_desc_sw128(smem_u32(b_smem), 16, 1024);
    const uint64_t bdesc_mn = make_smem_desc_sw128(smem_u32(b_smem), kBlockBytes, 1024);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        for (int o = 0; o < 8; ++o) {
          const uint32_t op = uint32_t(pattern >> (8 * o)) & 0xffu;
          if (!(op & 1u)) break;
          const int n = 64 << ((op >> 2) & 3u);
          const uint32_t mn = (op >> 4) & 1u, region = (op >> 5) & 3u;
          const uint32_t d = tmem_base + ((op >> 7) ? 224u : 0u);  // accumulators at columns 0 / 224, A at 480
          const uint32_t idesc = make_idesc_bf16_f32(128, n, 0, int(mn));
          const uint64_t bd = (mn ? bdesc_mn : bdesc_k) + uint64_t(region) * (kBlockBytes >> 4);
          for (int k = 0; k < group; ++k) {
            const uint64_t boff = mn ? uint64_t((k & 7) * 128) : uint64_t((k & 3) * 2);
            if (op & 2u) umma_ts_f16(d, tmem_base + 480u + uint32_t(k & 3) * 8u, bd + boff, idesc, 1);
            else         umma_ss_f16(d, adesc + uint64_t((k & 3) * 2), bd + boff, idesc, 1);
          }
        }
      }
      umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- compile-time MMA patterns (fully unrolled issue path, like the production kernels): one iteration = 4 k-steps.
//  0: SS128                     1: TS64 mn                   2: SS128 + TS64 k (same B tile)   3: SS128 + TS64 k (other B)
//  4: SS128 + TS64 mn (other B) 5: 4 x SS128 then 4 x TS64 mn 6: TS256 mn   7: TS128 mn   8: SS128 + 2 x TS64 mn
//  9: SS256                    10: SS256 + TS64 mn          11: SS256 + 2 x TS64 mn          12: SS128 + TS128 mn
// 13: TS128 k (A=Q-like) + TS64 mn                          14: SS64 + TS64 mn
template <int PAT>
__global__ void __launch_bounds__(128, 1) umma_bench2_kernel(int iters, unsigned long long* __restrict__ out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5;
  uintptr_t base = (reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~uintptr_t(1023u);
  uint8_t* a_smem = reinterpret_cast<uint8_t*>(base);  // 16 KiB
  uint8_t* b_smem = a_smem + kBlockBytes;              // regions 0..3, 16 KiB each
  uint64_t* done = reinterpret_cast<uint64_t*>(b_smem + 4 * kBlockBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  if (threadIdx.x == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  for (uint32_t i = threadIdx.x; i < 5 * kBlockBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_smem)[i] = 0;
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *tmem_slot;
  if (warp == 1) {
    constexpr uint32_t i_s64 = make_idesc_bf16_f32(128, 64, 0, 0), i_s128 = make_idesc_bf16_f32(128, 128, 0, 0),
                       i_s256 = make_idesc_bf16_f32(128, 256, 0, 0), i_o64 = make_idesc_bf16_f32(128, 64, 0, 1),
                       i_o128 = make_idesc_bf16_f32(128, 128, 0, 1), i_o256 = make_idesc_bf16_f32(128, 256, 0, 1);
    const uint64_t ad = make_smem_desc_sw128(smem_u32(a_smem), 16, 1024);
    const uint64_t bk = make_smem_desc_sw128(smem_u32(b_smem), 16, 1024);                      // K-major, region 0
    const uint64_t bk2 = bk + 2 * (kBlockBytes >> 4);                                            // K-major, region 2
    const uint64_t bm = make_smem_desc_sw128(smem_u32(b_smem), kBlockBytes, 1024);             // MN-major, region 0..
    const uint64_t bm2 = bm + 2 * (kBlockBytes >> 4);                                            // MN-major, region 2
    const uint32_t dS = tb, dO = tb + 256u, aP = tb + 448u;
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        if (PAT == 5) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_f16(dS, ad + uint64_t(k * 2), bk + uint64_t(k * 2), i_s128, 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts_f16(dO, aP + uint32_t(k) * 8u, bm2 + uint64_t(k * 128), i_o64, 1);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ka = uint64_t(k * 2), km = uint64_t(k * 128);
            const uint32_t pa = aP + uint32_t(k) * 8u;
            if (PAT == 0 || PAT == 2 || PAT == 3 || PAT == 4 || PAT == 8 || PAT == 12) umma_ss_f16(dS, ad + ka, bk + ka, i_s128, 1);
            if (PAT == 9 || PAT == 10 || PAT == 11) umma_ss_f16(dS, ad + ka, bk + ka, i_s256, 1);
            if (PAT == 14) umma_ss_f16(dS, ad + ka, bk + ka, i_s64, 1);
            if (PAT == 13) umma_ts_f16(dS, pa, bk + ka, i_s128, 1);
            if (PAT == 2) umma_ts_f16(dO, pa, bk + ka, i_s64, 1);
            if (PAT == 3) umma_ts_f16(dO, pa, bk2 + ka, i_s64, 1);
            if (PAT == 1 || PAT == 4 || PAT == 8 || PAT == 10 || PAT == 11 || PAT == 13 || PAT == 14) umma_ts_f16(dO, pa, bm2 + km, i_o64, 1);
            if (PAT == 8 || PAT == 11) umma_ts_f16(dO + 64u, pa, bm2 + km + (kBlockBytes >> 4), i_o64, 1);
            if (PAT == 6) umma_ts_f16(dO, pa, bm + km, i_o256, 0);
            if (PAT == 7 || PAT == 12) umma_ts_f16(dO, pa, bm2 + km, i_o128, 1);
          }
        }
      }
      umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int PAT>
int launch_bench2(int iters, unsigned long long* out, int grid, cudaStream_t stream) {
  const size_t smem = 1024 + 5 * kBlockBytes + 128;
  SPN_CUDA_CHECK(cudaFuncSetAttribute(umma_bench2_kernel<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  umma_bench2_kernel<PAT><<<grid, 128, smem, stream>>>(iters, out);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

// ---- DSMEM store bandwidth probe: clusters of 4; ranks 0/1 write `iters` x 32 KiB to rank 2 (and rank 3 when
// ndst == 2) with st.shared::cluster.v4, the way an S-CTA would ship bf16 P tiles to O-CTAs.
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1)
dsmem_bench_kernel(int iters, int ndst, unsigned long long* __restrict__ out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t rank = cluster_ctarank();
  uint8_t* buf = smem_raw;  // 2 x 32 KiB landing zones (one per writer) in the receivers
  cluster_sync_all();
  if (rank < 2) {
    const uint32_t local = smem_u32(buf) + rank * 32768u + threadIdx.x * 256u;
    const uint32_t d0 = mapa_shared(local, 2), d1 = mapa_shared(local, 3);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        st_cluster_v4(d0 + v * 16, it, v, rank, 1);
        if (ndst == 2) st_cluster_v4(d1 + v * 16, it, v, rank, 2);
      }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  } else if (threadIdx.x == 0) {
    out_cycles[blockIdx.x] = 0;
  }
  cluster_sync_all();  // no CTA may exit while its shared memory can still be written remotely
}

}  // namespace

int debug_dsmem_bench(int iters, int ndst, unsigned long long* out_cycles, int grid, cudaStream_t stream) {
  int rc;
  if ((rc = require_sm100()) != SPN_OK) return rc;
  SPN_REQUIRE(grid % 4 == 0 && grid > 0, SPN_ERR_BAD_SHAPE, "dsmem bench grid must be a multiple of 4");
  const size_t smem = 65536;
  SPN_CUDA_CHECK(cudaFuncSetAttribute(dsmem_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  dsmem_bench_kernel<<<grid, 128, smem, stream>>>(iters, ndst, out_cycles);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int debug_umma_bench(const void* b_src, long long b_rows, int N, int iters, int mode, int tma_on,
                     unsigned long long* out_cycles, int grid, cudaStream_t stream) {
  // (b_src, b_rows, N) are unused since the probe became pattern-driven: mode/tma_on carry the low/high 32 bits of the
  // 8-op pattern, N the group length (k-steps per op before the next op is issued)
  (void)b_src; (void)b_rows;
  int rc;
  if ((rc = require_sm100()) != SPN_OK) return rc;
  if (N == 0) {  // compile-time pattern `mode`
    switch (mode) {
#define SPN_B2(P) case P: return launch_bench2<P>(iters, out_cycles, grid, stream);
      SPN_B2(0) SPN_B2(1) SPN_B2(2) SPN_B2(3) SPN_B2(4) SPN_B2(5) SPN_B2(6) SPN_B2(7) SPN_B2(8) SPN_B2(9) SPN_B2(10)
      SPN_B2(11) SPN_B2(12) SPN_B2(13) SPN_B2(14)
#undef SPN_B2
      default: return fail(SPN_ERR_BAD_ARG, "unknown bench pattern %d", mode);
    }
  }
  const unsigned long long pattern = (unsigned long long)(uint32_t)mode | ((unsigned long long)(uint32_t)tma_on << 32);
  SPN_REQUIRE(N >= 1 && N <= 64, SPN_ERR_BAD_SHAPE, "bench group length must be in [1, 64]");
  const size_t smem = 1024 + 5 * kBlockBytes + 128;
  SPN_CUDA_CHECK(cudaFuncSetAttribute(umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  umma_bench_kernel<<<grid, 128, smem, stream>>>(pattern, iters, N, out_cycles);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int debug_umma_gemm(const void* a, const void* b, float* c, int N, int K, int mode, cudaStream_t stream) {
  SPN_REQUIRE(N == 64 || N == 128 || (N == 256 && (mode & 4)), SPN_ERR_BAD_SHAPE, "selftest N must be 64 or 128, or 256 with the wide MN-major mode (got %d)", N);
  SPN_REQUIRE(K % 64 == 0 && K >= 64 && K <= 256, SPN_ERR_BAD_SHAPE, "selftest K must be 64..256, multiple of 64 (got %d)", K);
  const bool b_mn = mode & 2;
  SPN_REQUIRE(!b_mn || K == 128, SPN_ERR_BAD_SHAPE, "MN-major selftest needs K == 128 (got %d)", K);
  int rc;
  if ((rc = require_sm100()) != SPN_OK) return rc;
  CUtensorMap ta, tb;
  if ((rc = make_tmap_bf16_2d(&ta, a, 128, uint64_t(K), 128)) != SPN_OK) return rc;
  if (b_mn) rc = make_tmap_bf16_2d(&tb, b, uint64_t(K), uint64_t(N), 128);
  else      rc = make_tmap_bf16_2d(&tb, b, uint64_t(N), uint64_t(K), uint32_t(N));
  if (rc != SPN_OK) return rc;
  const size_t smem = 1024 + 8 * kBlockBytes + 64;
  SPN_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  umma_selftest_kernel<<<1, 128, smem, stream>>>(ta, tb, static_cast<const uint32_t*>(a), c, N, K, mode);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
