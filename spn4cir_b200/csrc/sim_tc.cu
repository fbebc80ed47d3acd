// Retrieval scoring on the tensor cores: exact ranks / top-K from a tcgen05 score sweep + fp32 rescoring.
// Same contract as sim_rank.cu (`distances = 1 - q @ g.T; argsort` of clip4cir/validate.py:31-32,126-127 and the
// max-over-32-tokens similarity of blip2_qformer_cir_align_prompt.py:311-319): descending score, ties -> lower index,
// and the scores that leave this file are the SAME fp32 bits sim_rank.cu produces (ascending-k FMA chain).
//
// Three kernels per call:
//   1. split_rows_kernel      fp32 rows -> bf16 hi + bf16 lo (x = hi + lo + O(2^-18 |x|)), max row norm of the gallery
//   2. sim_tc_gemm_kernel     persistent tcgen05 GEMM, S~ = Qh Gh^T + Qh Gl^T + Ql Gh^T accumulated in one TMEM tile
//                             (fp32), 128 x 128 tiles, TMA-fed 3-stage ring, double-buffered accumulators, TMA-store
//                             epilogue (T = 32: the max over each entry's 32 token columns is taken in registers)
//                             -> S~ [Q, G] fp32 (L2-resident at validation sizes; |S~ - S| <= eps = 4e-5 |q| max|g|)
//   3. row_topk_kernel / row_rank_kernel   one warp per query row over S~:
//        top-K: shared-memory histogram of the row -> bin of the 64th best -> the <= 128 entries at or above it are
//               compacted and bitonic-sorted by (S~, index) -> window of 64.  Entries closer than 2 eps to a neighbour
//               (S~ cannot order them) are rescored with the exact fp32 chain -- all K returned entries when exact
//               values are requested -- and a 64-element bitonic sort on (cluster, exact score, index) finishes.  The
//               row is certified when (64th S~ + 2 eps) < (K-th S~); otherwise (plateaus wider than the window) it
//               takes the general path: all 64 rescored and a second scan rescoring every outside entry within eps of
//               the K-th exact score.  K <= 56.
//        rank : the target's exact score (one chain, or the caller's value when the gallery is sharded); entries
//               farther than eps from it are counted from S~, the others are rescored exactly.
//      The decision rules are restated and fuzzed on the CPU in tests/test_tc_decision_model.py.
// The tensor-core sweep decides nothing by itself: it only prunes which entries need the exact chain.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "bank_ce.cuh"
#include "ptx.cuh"
#include "sim_rank.cuh"
#include "topk_list.cuh"

namespace spn {
namespace {

constexpr int kTcThreads = 192;      // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int kTcStages = 3;         // ring stages of {Qh, Ql, Gh, Gl} k-blocks (4 x 16 KiB)
constexpr uint32_t kStageBytes = 4 * kBlockBytes;
constexpr int kSlotBytes = 4096;     // [32 rows x 32 fp32] staging slot of the TMA-store epilogue
constexpr float kEpsRel = 4e-5f;     // bound used for |S~ - S| / (|q| max|g|): 3 * 2^-18 from the dropped split terms
                                     // (1.1e-5) plus fp32 accumulation noise, with margin

// ------------------------------------------------------------------------------------------------ split prologue
// one warp per row: hi = bf16(x), lo = bf16(x - hi); optional max over rows of |row| (atomicMax on the float bits)
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ src, long long rows, int D, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo, float* __restrict__ maxnorm) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* s4 = reinterpret_cast<const float4*>(src + row * D);
  uint2* h2 = reinterpret_cast<uint2*>(hi + row * D);
  uint2* l2 = reinterpret_cast<uint2*>(lo + row * D);
  float ss = 0.f;
  for (int c = lane; c < D / 4; c += 32) {
    const float4 x = s4[c];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x.x), h1 = __float2bfloat16_rn(x.y), h2v = __float2bfloat16_rn(x.z),
                        h3 = __float2bfloat16_rn(x.w);
    const float r0 = x.x - __bfloat162float(h0), r1 = x.y - __bfloat162float(h1), r2 = x.z - __bfloat162float(h2v),
                r3 = x.w - __bfloat162float(h3);
    uint2 ho, lw;
    ho.x = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
    ho.y = uint32_t(__bfloat16_as_ushort(h2v)) | (uint32_t(__bfloat16_as_ushort(h3)) << 16);
    lw.x = pack_bf16x2(r0, r1);
    lw.y = pack_bf16x2(r2, r3);
    h2[c] = ho;
    l2[c] = lw;
    ss = fmaf(x.x, x.x, ss);
    ss = fmaf(x.y, x.y, ss);
    ss = fmaf(x.z, x.z, ss);
    ss = fmaf(x.w, x.w, ss);
  }
  if (maxnorm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
    if (lane == 0) atomicMax(reinterpret_cast<int*>(maxnorm), __float_as_int(sqrtf(ss)));  // non-negative floats order as ints
  }
}

// ------------------------------------------------------------------------------------------------ tcgen05 sweep
struct TcGemmParams {
  int q_row0;     // first query row of this block (coordinate into the split query maps)
  int qn;         // queries in this block
  int G;          // gallery entries
  int Gp;         // score row pitch (G rounded up to 4)
  int kblocks;    // D / 64
  int q_tiles, g_tiles;
  int nterms;     // 3: hi*hi + hi*lo + lo*hi, 1: hi*hi only
  float* scores;  // [qn, Gp]
  // fused rank epilogue (RANK = true, T = 1): no S~ matrix -- each epilogue thread counts its query's entries that
  // certainly sort before the target (S~ above the decision band), records the undecided ones (inside the band) as a 128-bit mask per
  // (query, gallery tile) and the S~ of the few columns that need individual treatment (the CIRR subset members)
  const float* thr;      // [qn, 3]  decision band (ts - eps, ts + eps) and ts
  const int* cols;       // [qn, M1] local columns to extract (members; -1 = none)
  float* col_s;          // [qn, M1] their S~ (pre-set to -inf)
  int M1;
  const int* skip;       // [qn, 2]  local columns excluded from the count: the removed reference image, the target
  uint4* umask;          // [qn, g_tiles]          (1-CTA sweep: one 128-bit mask per 128-column tile)
  uint32_t* umask32;     // [qn, n_chunks]         (pair sweep: one 32-bit mask per 32-column chunk)
  int n_chunks;
  int* rank;             // [qn] += certain count
  // pair sweep (sim_tc_gemm2_kernel): 256 queries x tile_n gallery rows per cluster tile
  int tile_n;            // multiple of 32, <= 256
  int q_tiles2, g_tiles2;
};

template <bool T32, bool RANK = false>
__global__ void __launch_bounds__(kTcThreads, 1)
sim_tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_qh, const __grid_constant__ CUtensorMap tmap_ql,
                   const __grid_constant__ CUtensorMap tmap_gh, const __grid_constant__ CUtensorMap tmap_gl,
                   const __grid_constant__ CUtensorMap tmap_s, const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~uintptr_t(1023u));
  uint8_t* ring = base;                                        // kTcStages x 64 KiB
  uint8_t* slots = ring + size_t(kTcStages) * kStageBytes;     // 4 warps x 2 slots x 4 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + 8 * kSlotBytes);
  uint64_t* full = bars;                    // [kTcStages] TMA -> MMA
  uint64_t* empty = bars + kTcStages;       // [kTcStages] MMA -> TMA
  uint64_t* acc_full = bars + 2 * kTcStages;      // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * kTcStages + 2; // [2] epilogue -> MMA (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_qh);
    tma_prefetch_desc(&tmap_ql);
    tma_prefetch_desc(&tmap_gh);
    tma_prefetch_desc(&tmap_gl);
    tma_prefetch_desc(&tmap_s);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = p.q_tiles * p.g_tiles;
  const bool three = p.nterms == 3;

  if (warp == 0) {
    // ---------------- TMA producer
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int qt = tile / p.g_tiles, gt = tile % p.g_tiles;
      const int qrow = p.q_row0 + qt * kBM, grow = gt * kBN;
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        if (elect_one()) {
          uint8_t* st = ring + size_t(s) * kStageBytes;
          mbar_arrive_expect_tx(&full[s], three ? kStageBytes : 2 * kBlockBytes);
          tma_load_2d(st, &tmap_qh, &full[s], kb * kBK, qrow);
          tma_load_2d(st + 2 * kBlockBytes, &tmap_gh, &full[s], kb * kBK, grow);
          if (three) {
            tma_load_2d(st + kBlockBytes, &tmap_ql, &full[s], kb * kBK, qrow);
            tma_load_2d(st + 3 * kBlockBytes, &tmap_gl, &full[s], kb * kBK, grow);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    const uint32_t idesc = make_idesc_bf16_f32(kBM, kBN, 0, 0);
    uint32_t it = 0, j = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const uint32_t b = j & 1u, aph = (j >> 1) & 1u;
      mbar_wait(&acc_empty[b], aph ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + b * uint32_t(kBN);
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(ring + size_t(s) * kStageBytes);
          const uint64_t qh = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t ql = make_smem_desc_sw128(sa + kBlockBytes, 16, 1024);
          const uint64_t gh = make_smem_desc_sw128(sa + 2 * kBlockBytes, 16, 1024);
          const uint64_t gl = make_smem_desc_sw128(sa + 3 * kBlockBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t o = uint64_t(k * 2);
            if (three) {  // small terms first
              umma_ss_f16(d, qh + o, gl + o, idesc, (kb | k) != 0);
              umma_ss_f16(d, ql + o, gh + o, idesc, 1);
              umma_ss_f16(d, qh + o, gh + o, idesc, 1);
            } else {
              umma_ss_f16(d, qh + o, gh + o, idesc, (kb | k) != 0);
            }
          }
          umma_commit(&empty[s]);
          if (kb == p.kblocks - 1) umma_commit(&acc_full[b]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> (swizzled staging slot -> TMA store | max over token groups)
    const uint32_t quarter = warp & 3u;  // the TMEM lane quarter this warp may read
    const uint32_t t_row = tmem_base + ((quarter * 32u) << 16);
    uint8_t* my_slots = slots + size_t(warp - 2u) * 2 * kSlotBytes;
    uint32_t x[32];
    uint32_t j = 0, k = 0;
    int rk_cnt = 0, rk_row = 0, rk_s0 = -1, rk_s1 = -1;
    bool rk_ok = false;
    float rk_lo = 0.f, rk_hi = 0.f;
    uint32_t rk_mask[4] = {0u, 0u, 0u, 0u};
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
      const int qt = tile / p.g_tiles, gt = tile % p.g_tiles;
      const uint32_t b = j & 1u, aph = (j >> 1) & 1u;
      const int r0 = qt * kBM + int(quarter) * 32;  // first block-relative query row of this warp
      mbar_wait(&acc_full[b], aph);
      tc_fence_after();
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32b_x32(t_row + b * uint32_t(kBN) + uint32_t(c) * 32u, x);
        tmem_ld_wait();
        if (c == 3) {  // accumulator drained: the MMA warp may start the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[b]);
        }
        if (RANK) {
          // fused rank epilogue: thread == query row; count / mask this 32-column chunk against the row's two thresholds
          if (c == 0) {
            rk_cnt = 0;
            rk_row = r0 + int(lane);
            rk_ok = rk_row < p.qn;
            rk_lo = rk_ok ? p.thr[3 * rk_row] : INFINITY;
            rk_hi = rk_ok ? p.thr[3 * rk_row + 1] : INFINITY;
            rk_s0 = rk_ok ? p.skip[2 * rk_row] : -1;
            rk_s1 = rk_ok ? p.skip[2 * rk_row + 1] : -1;
          }
          const int col0 = gt * kBN + c * 32;
          uint32_t m = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = __uint_as_float(x[i]);
            const int col = col0 + i;
            const bool ok = col < p.G && col != rk_s0 && col != rk_s1;
            const bool before = ok && a > rk_hi;
            const bool unsure = ok && !before && !(a < rk_lo);
            rk_cnt += before ? 1 : 0;
            m |= unsure ? (1u << i) : 0u;
          }
          rk_mask[c] = m;
          if (rk_ok) {  // columns whose S~ the resolve kernel wants individually (rare: a handful per query over all tiles)
            for (int mm = 0; mm < p.M1; ++mm) {
              const unsigned d = unsigned(p.cols[size_t(rk_row) * p.M1 + mm] - col0);
              if (d < 32u) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (unsigned(i) == d) v = __uint_as_float(x[i]);
                p.col_s[size_t(rk_row) * p.M1 + mm] = v;
              }
            }
          }
          if (c == 3 && rk_ok) {
            if (rk_cnt) atomicAdd(&p.rank[rk_row], rk_cnt);
            p.umask[size_t(rk_row) * p.g_tiles + gt] = make_uint4(rk_mask[0], rk_mask[1], rk_mask[2], rk_mask[3]);
          }
        } else
        if (T32) {
          float m = __uint_as_float(x[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(x[i]));
          mx[c] = m;
        } else {
          const int col0 = gt * kBN + c * 32;
          if (r0 < p.qn && col0 < p.Gp) {  // warp-uniform
            uint8_t* slot = my_slots + size_t(k & 1u) * kSlotBytes;
            if (k >= 2) {
              if (lane == 0) bulk_wait_group_read<1>();
              __syncwarp();
            }
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<uint4*>(slot + lane * 128u + ((uint32_t(v) ^ (lane & 7u)) << 4)) =
                  make_uint4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(slot, &tmap_s, col0, r0);
              bulk_commit_group();
            }
            ++k;
          }
        }
      }
      if (T32) {
        const int row = r0 + int(lane);
        if (row < p.qn) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int e = gt * 4 + c;
            if (e < p.G) p.scores[size_t(row) * p.Gp + e] = mx[c];
          }
        }
      }
    }
    if (!T32 && !RANK) {
      if (lane == 0) bulk_wait_group<0>();
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ pair sweep
// The same sweep on CTA pairs (cta_group::2): one cluster tile = 256 queries x tile_n gallery rows.  Each CTA stages its own
// 128 query rows (hi, lo) and HALF of the tile's gallery rows (hi, lo); rank 0 issues M256 x N(tile_n) x K16 MMAs for the
// pair, so every operand byte fetched from L2 feeds twice the output of the 1-CTA tile (the 1-CTA sweep is L2 -> SM bound:
// 389 MB of operand traffic at CIRR size) and the fixed cost of an SS instruction (the 4 KiB A fetch, ~43 cycles) is
// amortised over up to 256 columns instead of 128.  tile_n is chosen per problem (multiple of 32, <= 256) to minimise
// ceil(tiles / clusters) x instruction time -- at CIRR size 192 (204 tiles on 74 clusters: 2.76 -> 3 rounds).
template <bool T32, bool RANK>
__global__ void __launch_bounds__(kTcThreads, 1)
sim_tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_qh, const __grid_constant__ CUtensorMap tmap_ql,
                    const __grid_constant__ CUtensorMap tmap_gh, const __grid_constant__ CUtensorMap tmap_gl,
                    const __grid_constant__ CUtensorMap tmap_s, const TcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~uintptr_t(1023u));
  uint8_t* ring = base;                                        // kTcStages x 64 KiB: Qh | Ql | Gh half | Gl half
  uint8_t* slots = ring + size_t(kTcStages) * kStageBytes;     // 4 warps x 2 slots x 4 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + 8 * kSlotBytes);
  uint64_t* full = bars;                    // [kTcStages] TMA (both CTAs) -> MMA, rank 0's are used
  uint64_t* empty = bars + kTcStages;       // [kTcStages] MMA -> TMA, multicast to both CTAs
  uint64_t* acc_full = bars + 2 * kTcStages;      // [2] MMA -> epilogue, multicast
  uint64_t* acc_empty = bars + 2 * kTcStages + 2; // [2] epilogue (4 warps x 2 CTAs) -> MMA, rank 0's are used
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);
  const uint32_t crank = cluster_ctarank();
  const int cluster_id = int(blockIdx.x >> 1), n_clusters = int(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_qh);
    tma_prefetch_desc(&tmap_ql);
    tma_prefetch_desc(&tmap_gh);
    tma_prefetch_desc(&tmap_gl);
    if (!T32 && !RANK) tma_prefetch_desc(&tmap_s);
  }
  if (warp == 1) {
    tmem_alloc_cg2(tmem_slot, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  cluster_sync_all();  // the peer's barriers must be initialised before anything is signalled remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = p.q_tiles2 * p.g_tiles2;
  const int half_n = p.tile_n >> 1;
  const uint32_t g_bytes = uint32_t(half_n) * 128u;  // one k-block of this CTA's gallery half

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs; completions are credited to rank 0's barrier)
    uint32_t it = 0;
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
      const int qt = tile / p.g_tiles2, gt = tile % p.g_tiles2;
      const int qrow = p.q_row0 + qt * 2 * kBM + int(crank) * kBM;
      const int grow = gt * p.tile_n + int(crank) * half_n;
      for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
        const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t fbar = mapa_shared(smem_u32(&full[s]), 0);
        if (elect_one()) {
          uint8_t* st = ring + size_t(s) * kStageBytes;
          if (crank == 0) mbar_arrive_expect_tx(&full[s], 2u * (2u * kBlockBytes + 2u * g_bytes));
          tma_load_2d_cg2(st, &tmap_qh, fbar, kb * kBK, qrow);
          tma_load_2d_cg2(st + kBlockBytes, &tmap_ql, fbar, kb * kBK, qrow);
          tma_load_2d_cg2(st + 2 * kBlockBytes, &tmap_gh, fbar, kb * kBK, grow);
          tma_load_2d_cg2(st + 3 * kBlockBytes, &tmap_gl, fbar, kb * kBK, grow);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (rank 0 issues for the pair)
    if (crank == 0) {
      const uint32_t idesc = make_idesc_bf16_f32(2 * kBM, p.tile_n, 0, 0);
      uint32_t it = 0, j = 0;
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++j) {
        const uint32_t b = j & 1u, aph = (j >> 1) & 1u;
        mbar_wait(&acc_empty[b], aph ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + b * 256u;
        for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
          const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1u;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(ring + size_t(s) * kStageBytes);
            const uint64_t qh = make_smem_desc_sw128(sa, 16, 1024);
            const uint64_t ql = make_smem_desc_sw128(sa + kBlockBytes, 16, 1024);
            const uint64_t gh = make_smem_desc_sw128(sa + 2 * kBlockBytes, 16, 1024);
            const uint64_t gl = make_smem_desc_sw128(sa + 3 * kBlockBytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t o = uint64_t(k * 2);
              umma_ss_f16_cg2(d, qh + o, gl + o, idesc, (kb | k) != 0);  // small terms first
              umma_ss_f16_cg2(d, ql + o, gh + o, idesc, 1);
              umma_ss_f16_cg2(d, qh + o, gh + o, idesc, 1);
            }
            umma_commit_cg2(&empty[s], 3);
            if (kb == p.kblocks - 1) umma_commit_cg2(&acc_full[b], 3);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue (both CTAs): TMEM -> registers -> (staging slot -> TMA store | token max | rank count)
    const uint32_t quarter = warp & 3u;
    const uint32_t t_row = tmem_base + ((quarter * 32u) << 16);
    uint8_t* my_slots = slots + size_t(warp - 2u) * 2 * kSlotBytes;
    const int nchunk = p.tile_n >> 5;
    uint32_t x[32];
    uint32_t j = 0, k = 0;
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++j) {
      const int qt = tile / p.g_tiles2, gt = tile % p.g_tiles2;
      const uint32_t b = j & 1u, aph = (j >> 1) & 1u;
      const int r0 = qt * 2 * kBM + int(crank) * kBM + int(quarter) * 32;  // first block-relative query row of this warp
      const int row = r0 + int(lane);
      const bool row_ok = row < p.qn;
      float rk_lo = INFINITY, rk_hi = INFINITY;
      int rk_s0 = -1, rk_s1 = -1, rk_cnt = 0;
      if (RANK && row_ok) {
        rk_lo = p.thr[3 * row];
        rk_hi = p.thr[3 * row + 1];
        rk_s0 = p.skip[2 * row];
        rk_s1 = p.skip[2 * row + 1];
      }
      mbar_wait(&acc_full[b], aph);
      tc_fence_after();
      for (int c = 0; c < nchunk; ++c) {
        tmem_ld_32x32b_x32(t_row + b * 256u + uint32_t(c) * 32u, x);
        tmem_ld_wait();
        if (c == nchunk - 1) {  // accumulator drained: the MMA warp may start the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&acc_empty[b]), 0));
        }
        const int col0 = gt * p.tile_n + c * 32;
        if (RANK) {
          uint32_t m = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = __uint_as_float(x[i]);
            const int col = col0 + i;
            const bool ok = col < p.G && col != rk_s0 && col != rk_s1;
            const bool before = ok && a > rk_hi;
            const bool unsure = ok && !before && !(a < rk_lo);
            rk_cnt += before ? 1 : 0;
            m |= unsure ? (1u << i) : 0u;
          }
          if (row_ok && col0 < p.G) {
            p.umask32[size_t(row) * p.n_chunks + (col0 >> 5)] = m;
            for (int mm = 0; mm < p.M1; ++mm) {  // columns whose S~ the resolve kernel wants individually (rare hits)
              const unsigned dd = unsigned(p.cols[size_t(row) * p.M1 + mm] - col0);
              if (dd < 32u) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (unsigned(i) == dd) v = __uint_as_float(x[i]);
                p.col_s[size_t(row) * p.M1 + mm] = v;
              }
            }
          }
        } else if (T32) {
          float m = __uint_as_float(x[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(x[i]));
          const int e = col0 >> 5;
          if (row_ok && e < p.G) p.scores[size_t(row) * p.Gp + e] = m;
        } else {
          if (r0 < p.qn && col0 < p.Gp) {  // warp-uniform
            uint8_t* slot = my_slots + size_t(k & 1u) * kSlotBytes;
            if (k >= 2) {
              if (lane == 0) bulk_wait_group_read<1>();
              __syncwarp();
            }
#pragma unroll
            for (int v = 0; v < 8; ++v)
              *reinterpret_cast<uint4*>(slot + lane * 128u + ((uint32_t(v) ^ (lane & 7u)) << 4)) =
                  make_uint4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(slot, &tmap_s, col0, r0);
              bulk_commit_group();
            }
            ++k;
          }
        }
      }
      if (RANK && row_ok && rk_cnt) atomicAdd(&p.rank[row], rk_cnt);
    }
    if (!T32 && !RANK) {
      if (lane == 0) bulk_wait_group<0>();
      __syncwarp();
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrivals can still land
  if (warp == 1) tmem_dealloc_cg2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ exact rescoring
// The reference chain of sim_rank.cu: one fp32 accumulator per score, fmaf in ascending k (max over T token rows).
constexpr int kRowWarps = 4;
constexpr int kChunkK = 32;
constexpr int kChainPitch = kChunkK + 4;                        // 16-byte aligned rows, float4 walk
constexpr int kFewW = 128;                                      // columns per step of the <= 8-row variant
constexpr int kFewPitch = kFewW + 4;                            // 8 rows x 4 banks: conflict-free float4 walk
constexpr int kChainFloats = 8 * kFewPitch + kFewW + 32;        // >= 32 * kChainPitch + kChunkK + 32: rows | q | row ids
constexpr int kBins = 1024;                                     // histogram bins of the top-K threshold search
constexpr int kCand = 128;                                      // candidates kept above the threshold bin
constexpr int kRowSmemFloats = kBins > kChainFloats ? kBins : kChainFloats;  // the two phases share the space

// Up to 32 chains at once, one per lane: returns the chain of q . g[row] for this lane's gallery row (`row` < 0: none).
// The warp stages 32-column chunks of the <= 32 gallery rows (16-byte loads: one instruction fetches 128 bytes of four
// rows) and of q in shared memory and every lane walks its own staged row; the loads of chunk c + 1 are in flight while
// chunk c is walked.
__device__ __forceinline__ float warp_chains(const float* __restrict__ qrow, const float* __restrict__ g, int D, int row,
                                             float* sm, int lane) {
  float acc = 0.f;
  float* sq = sm + 32 * kChainPitch;
  int* srows = reinterpret_cast<int*>(sq + kChunkK);
  const unsigned have = __ballot_sync(kFull, row >= 0);
  __syncwarp();
  srows[lane] = row;
  __syncwarp();
  const int sub = lane >> 3, c4 = (lane & 7) * 4;
  float4 t[8], tq;
  int my_rows[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) my_rows[h] = srows[h * 4 + sub];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (((have >> (h * 4)) & 0xfu) == 0) continue;  // warp-uniform
      const int rr = my_rows[h];
      t[h] = rr >= 0 ? *reinterpret_cast<const float4*>(g + size_t(rr) * D + k0 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (lane < 8) tq = *reinterpret_cast<const float4*>(qrow + k0 + c4);
  };
  load_chunk(0);
  for (int k0 = 0; k0 < D; k0 += kChunkK) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (((have >> (h * 4)) & 0xfu) == 0) continue;
      *reinterpret_cast<float4*>(sm + (h * 4 + sub) * kChainPitch + c4) = t[h];
    }
    if (lane < 8) *reinterpret_cast<float4*>(sq + c4) = tq;
    __syncwarp();
    if (k0 + kChunkK < D) load_chunk(k0 + kChunkK);
    const float4* mine4 = reinterpret_cast<const float4*>(sm + lane * kChainPitch);
    const float4* sq4 = reinterpret_cast<const float4*>(sq);
#pragma unroll
    for (int k = 0; k < kChunkK / 4; ++k) {
      const float4 a = sq4[k], b = mine4[k];
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc);
      acc = fmaf(a.w, b.w, acc);
    }
    __syncwarp();
  }
  return acc;
}

// Same chains for at most 8 rows, held by lanes 0..7: 128 columns per step (one 512-byte load instruction per row, nine
// loads in flight, a quarter of the dependent L2 round trips) -- the common case: a near-tie cluster, the target row, a
// few unsure entries.
__device__ __forceinline__ float warp_chains_few(const float* __restrict__ qrow, const float* __restrict__ g, int D, int row,
                                                 float* sm, int lane) {
  float acc = 0.f;
  float* sq = sm + 8 * kFewPitch;
  int* srows = reinterpret_cast<int*>(sq + kFewW);
  __syncwarp();
  srows[lane] = lane < 8 ? row : -1;
  __syncwarp();
  int rows8[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) rows8[u] = srows[u];
  float4 t[8], tq;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_step = [&](int k0) {
    const bool in = k0 + lane * 4 < D;
#pragma unroll
    for (int u = 0; u < 8; ++u)
      t[u] = (in && rows8[u] >= 0) ? *reinterpret_cast<const float4*>(g + size_t(rows8[u]) * D + k0 + lane * 4) : zero4;
    tq = in ? *reinterpret_cast<const float4*>(qrow + k0 + lane * 4) : zero4;
  };
  load_step(0);
  const float* mine = sm + (lane & 7) * kFewPitch;
  for (int k0 = 0; k0 < D; k0 += kFewW) {
#pragma unroll
    for (int u = 0; u < 8; ++u) *reinterpret_cast<float4*>(sm + u * kFewPitch + lane * 4) = t[u];
    *reinterpret_cast<float4*>(sq + lane * 4) = tq;
    __syncwarp();
    if (k0 + kFewW < D) load_step(k0 + kFewW);
    const int n4 = min(kFewW, D - k0) / 4;
    if (lane < 8 && row >= 0) {  // only the lanes that own a row walk it (a quarter of the shared-memory wavefronts)
      const float4* mine4 = reinterpret_cast<const float4*>(mine);
      const float4* sq4 = reinterpret_cast<const float4*>(sq);
#pragma unroll 8
      for (int k = 0; k < n4; ++k) {
        const float4 a = sq4[k], b = mine4[k];
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
    }
    __syncwarp();
  }
  return acc;
}

// Exact scores of gallery entries for the lanes that ask: (need0, e0) and (need1, e1) are two independent requests per
// lane (the two list positions a lane owns).  T = 1: the requests are compacted to the low lanes and served by one staged
// pass when they fit (<= 8: the few-row variant); T = 32: entry by entry, the 32 lanes take the 32 token rows and the
// maximum is reduced across the warp.
__device__ __forceinline__ void warp_exact2(const float* __restrict__ qrow, const float* __restrict__ g, int D, int T,
                                            bool need0, int e0, bool need1, int e1, float* sm, int lane, float& out0,
                                            float& out1) {
  out0 = out1 = -INFINITY;
  const unsigned m0 = __ballot_sync(kFull, need0), m1 = __ballot_sync(kFull, need1);
  const int n0 = __popc(m0), n1 = __popc(m1);
  if (n0 + n1 == 0) return;
  if (T == 1) {
    if (n0 + n1 <= 32) {
      // lane j serves request j: the j-th set bit of m0, then of m1
      const unsigned src0 = __fns(m0, 0, lane + 1), src1 = __fns(m1, 0, lane - n0 + 1);
      const int a = __shfl_sync(kFull, e0, src0 & 31u), b = __shfl_sync(kFull, e1, src1 & 31u);
      const int mine = lane < n0 ? a : (lane < n0 + n1 ? b : -1);
      const float acc = (n0 + n1 <= 8) ? warp_chains_few(qrow, g, D, mine, sm, lane) : warp_chains(qrow, g, D, mine, sm, lane);
      const float r0 = __shfl_sync(kFull, acc, __popc(m0 & ((1u << lane) - 1u)) & 31);
      const float r1 = __shfl_sync(kFull, acc, (n0 + __popc(m1 & ((1u << lane) - 1u))) & 31);
      if (need0) out0 = r0;
      if (need1) out1 = r1;
    } else {
      const float r0 = warp_chains(qrow, g, D, need0 ? e0 : -1, sm, lane);
      const float r1 = warp_chains(qrow, g, D, need1 ? e1 : -1, sm, lane);
      if (need0) out0 = r0;
      if (need1) out1 = r1;
    }
    return;
  }
  for (int pass = 0; pass < 2; ++pass) {
    unsigned mask = pass ? m1 : m0;
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const int ee = __shfl_sync(kFull, pass ? e1 : e0, src);
      float v = warp_chains(qrow, g, D, ee * T + lane, sm, lane);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
      if (lane == src) {
        if (pass) out1 = v;
        else out0 = v;
      }
    }
  }
}
__device__ __forceinline__ float warp_exact(const float* __restrict__ qrow, const float* __restrict__ g, int D, int T,
                                            bool need, int e, float* sm, int lane) {
  float a, b;
  warp_exact2(qrow, g, D, T, need, e, false, 0, sm, lane, a, b);
  return a;
}

// per-lane gather version for the rare paths (second scan, entries within eps of the target score)
__device__ __forceinline__ float exact_one(const float* __restrict__ q, const float* __restrict__ g, int D, int T,
                                           int e) {
  float best = -INFINITY;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  for (int tok = 0; tok < T; ++tok) {
    const float4* a4 = reinterpret_cast<const float4*>(g + (size_t(e) * T + tok) * D);
    float acc = 0.f;
#pragma unroll 4
    for (int c = 0; c < D / 4; ++c) {
      const float4 qq = q4[c], aa = a4[c];
      acc = fmaf(qq.x, aa.x, acc);
      acc = fmaf(qq.y, aa.y, acc);
      acc = fmaf(qq.z, aa.z, acc);
      acc = fmaf(qq.w, aa.w, acc);
    }
    best = fmaxf(best, acc);
  }
  return best;
}

struct RowParams {
  const float* q;       // fp32 [Q, D] (row 0 = first row of this block)
  const float* g;       // fp32 [G*T, D]
  const float* scores;  // S~ [qn, Gp]
  const float* gmax;    // max gallery row norm
  int qn, G, Gp, D, T, K, col_base;
  int exact_values;     // top-K: out_val = exact fp32 chain (else S~; the ORDER is exact either way)
  const int* exclude;
  // top-K
  float* out_val;
  int* out_idx;
  int* n_second;        // rows that left the fast path (statistics)
  // rank
  const int* target;
  const float* tscore;
  const int* group;
  int M;
  int* rank;
  int* group_rank;
};

__device__ __forceinline__ float row_eps(const float* __restrict__ qrow, int D, float gmax, int lane) {
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) ss = fmaf(qrow[c], qrow[c], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
  return kEpsRel * sqrtf(ss) * gmax;
}

__device__ __forceinline__ void write_topk(const RowParams& p, int row, const float* srow, float x0, int j0, float x1,
                                           int j1, int lane) {
  const size_t o = size_t(row) * p.K;
  if (lane < p.K) {
    const bool real = j0 != INT_MAX;
    p.out_val[o + lane] = real ? (p.exact_values ? x0 : srow[j0 - p.col_base]) : -INFINITY;
    p.out_idx[o + lane] = real ? j0 : -1;
  }
  if (lane + 32 < p.K) {
    const bool real = j1 != INT_MAX;
    p.out_val[o + lane + 32] = real ? (p.exact_values ? x1 : srow[j1 - p.col_base]) : -INFINITY;
    p.out_idx[o + lane + 32] = real ? j1 : -1;
  }
}

// General path of a row (plateaus wider than the candidate window, lanes whose private list overflowed): shuffle-
// insertion top-64 by S~, all 64 rescored, and a second scan over everything within eps of the K-th exact score.
__device__ __noinline__ void row_topk_general(const RowParams& p, int row, const float* srow, const float* qrow, int ex,
                                              float eps, int lane) {
  float v0 = -INFINITY, v1 = -INFINITY;
  int i0 = INT_MAX, i1 = INT_MAX;
  for (int g0 = 0; g0 < p.G; g0 += 128) {
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + u * 32 + lane;
      a[u] = g < p.G ? srow[g] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + u * 32 + lane;
      const int idx = p.col_base + g;
      topk_offer(a[u], idx, g < p.G && idx != ex, 64, v0, i0, v1, i1, lane);
    }
  }
  const float a_last = __shfl_sync(kFull, v1, 31);
  const int i_last = __shfl_sync(kFull, i1, 31);
  const float e0 = i0 != INT_MAX ? exact_one(qrow, p.g, p.D, p.T, i0 - p.col_base) : -INFINITY;
  const float e1 = i1 != INT_MAX ? exact_one(qrow, p.g, p.D, p.T, i1 - p.col_base) : -INFINITY;
  __syncwarp();
  float x0 = -INFINITY, x1 = -INFINITY;
  int j0 = INT_MAX, j1 = INT_MAX;
  topk_offer(e0, i0, i0 != INT_MAX, 64, x0, j0, x1, j1, lane);
  topk_offer(e1, i1, i1 != INT_MAX, 64, x0, j0, x1, j1, lane);
  const int K = p.K;
  float e_k = (K - 1 < 32) ? __shfl_sync(kFull, x0, K - 1) : __shfl_sync(kFull, x1, K - 33);
  if (i_last != INT_MAX && !(a_last + eps < e_k)) {  // entries outside the window may still belong to the top K
    for (int g0 = 0; g0 < p.G; g0 += 32) {
      const int g = g0 + lane;
      const int idx = p.col_base + g;
      const float a = g < p.G ? srow[g] : -INFINITY;
      const bool cand = g < p.G && idx != ex && sorts_before(a_last, i_last, a, idx) && !(a + eps < e_k);
      if (__ballot_sync(kFull, cand)) {
        float e = -INFINITY;
        if (cand) e = exact_one(qrow, p.g, p.D, p.T, g);
        __syncwarp();
        topk_offer(e, idx, cand, K, x0, j0, x1, j1, lane);
        e_k = (K - 1 < 32) ? __shfl_sync(kFull, x0, K - 1) : __shfl_sync(kFull, x1, K - 33);
      }
    }
  }
  write_topk(p, row, srow, x0, j0, x1, j1, lane);
}

__device__ __forceinline__ bool key_before(int ca, float va, int ia, int cb, float vb, int ib) {
  return ca < cb || (ca == cb && sorts_before(va, ia, vb, ib));
}

// one warp per query row.  Fast path: a 1024-bin histogram of the row (linear in the score over [-|q| max|g|, +|q| max|g|],
// shared-memory atomics) gives the bin holding the 64th best entry; a second pass compacts the entries at or above that
// bin (<= 128, normally ~70) which one 128-element bitonic sort orders by (S~, index); exact chains run only where S~
// cannot decide the order (entries closer than 2 eps to a neighbour; all of the first K when exact values are
// requested), then a 64-element bitonic sort on (cluster, exact score, index).
__global__ void __launch_bounds__(kRowWarps * 32, 6)
row_topk_kernel(const RowParams p) {
  __shared__ float row_sm[kRowWarps][kRowSmemFloats];
  __shared__ float cand_v[kRowWarps][kCand];
  __shared__ int cand_i[kRowWarps][kCand];
  const int wib = threadIdx.x >> 5;
  const int row = blockIdx.x * kRowWarps + wib;
  const int lane = threadIdx.x & 31;
  if (row >= p.qn) return;
  const float* srow = p.scores + size_t(row) * p.Gp;
  const float* qrow = p.q + size_t(row) * p.D;
  const int ex = p.exclude ? p.exclude[row] : -1;
  const float eps = row_eps(qrow, p.D, *p.gmax, lane);
  const int K = p.K;

  // ---- pass 1: histogram
  unsigned* hist = reinterpret_cast<unsigned*>(row_sm[wib]);
#pragma unroll
  for (int b = 0; b < kBins / 32; ++b) hist[b * 32 + lane] = 0u;
  __syncwarp();
  const float bound = fmaxf(eps * (1.001f / kEpsRel), 1e-30f);  // >= |q| max|g| >= |S~|
  const float bscale = float(kBins / 2) / bound;
  auto bin_of = [&](float a) {
    const int b = __float2int_rd((a + bound) * bscale);
    return min(kBins - 1, max(0, b));
  };
  for (int g0 = 0; g0 < p.G; g0 += 256) {
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * 32 + lane;
      a[u] = g < p.G ? srow[g] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * 32 + lane;
      if (g < p.G && p.col_base + g != ex) atomicAdd(&hist[bin_of(a[u])], 1u);
    }
  }
  __syncwarp();
  // ---- threshold bin: lane L owns bins [32 L, 32 L + 32); walk down from the top until 64 entries are covered
  unsigned mine = 0;
#pragma unroll
  for (int b = 0; b < 32; ++b) mine += hist[lane * 32 + ((b + lane) & 31)];  // staggered: conflict-free
  unsigned above = mine;  // inclusive suffix sum over lanes >= this one
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_down_sync(kFull, above, o);
    if (lane + o < 32) above += t;
  }
  const unsigned total = __shfl_sync(kFull, above, 0);
  int tb = 0;
  unsigned n_cand = total;
  if (total > 64u) {
    const unsigned owner_mask = __ballot_sync(kFull, above >= 64u && above - mine < 64u);
    const int owner = __ffs(owner_mask) - 1;
    if (lane == owner) {
      unsigned c = above - mine;
      int b = 31;
      for (; b > 0; --b) {
        c += hist[lane * 32 + b];
        if (c >= 64u) break;
      }
      if (b == 0) c += hist[lane * 32];
      tb = lane * 32 + b;
      n_cand = c;
    }
    tb = __shfl_sync(kFull, tb, owner);
    n_cand = __shfl_sync(kFull, n_cand, owner);
  }
  if (n_cand > unsigned(kCand)) {  // a plateau wider than the candidate buffer
    if (lane == 0 && p.n_second) atomicAdd(p.n_second, 1);
    row_topk_general(p, row, srow, qrow, ex, eps, lane);
    return;
  }
  // ---- pass 2: compact the entries at or above the threshold bin
  unsigned fill = 0;
  for (int g0 = 0; g0 < p.G; g0 += 256) {
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * 32 + lane;
      a[u] = g < p.G ? srow[g] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * 32 + lane;
      const int idx = p.col_base + g;
      const bool take = g < p.G && idx != ex && bin_of(a[u]) >= tb;
      const unsigned m = __ballot_sync(kFull, take);
      if (take) {
        const unsigned pos = fill + __popc(m & ((1u << lane) - 1u));
        cand_v[wib][pos] = a[u];
        cand_i[wib][pos] = idx;
      }
      fill += __popc(m);
    }
  }
  __syncwarp();
  // ---- 128-element bitonic sort on (S~, index), position = lane + 32 r, best first
  float cv[4];
  int ci[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const unsigned pos = lane + 32 * r;
    cv[r] = pos < fill ? cand_v[wib][pos] : -INFINITY;
    ci[r] = pos < fill ? cand_i[wib][pos] : INT_MAX;
  }
#pragma unroll
  for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {  // partner register r ^ (j / 32) of the same lane
        const int rj = j >> 5;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if ((r & rj) == 0) {
            const int o = r | rj;
            const bool up = ((r * 32) & k) == 0;  // k = 64: r < 2 up, else down; k = 128: up
            const bool r_better = sorts_before(cv[r], ci[r], cv[o], ci[o]);
            if (r_better != up) {
              const float tv = cv[r]; cv[r] = cv[o]; cv[o] = tv;
              const int ti = ci[r]; ci[r] = ci[o]; ci[o] = ti;
            }
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool up = ((lane + 32 * r) & k) == 0;
          const float tv = __shfl_xor_sync(kFull, cv[r], j);
          const int ti = __shfl_xor_sync(kFull, ci[r], j);
          const bool mine_better = sorts_before(cv[r], ci[r], tv, ti);
          if (mine_better != (lower == up)) {
            cv[r] = tv;
            ci[r] = ti;
          }
        }
      }
    }
  }
  // the window: positions 0..63 are registers 0 and 1 (the list layout used below)
  float v0 = cv[0], v1 = cv[1];
  int i0 = ci[0], i1 = ci[1];
  const float a_last = __shfl_sync(kFull, v1, 31);
  const int i_last = __shfl_sync(kFull, i1, 31);
  const float a_k1 = (K - 1 < 32) ? __shfl_sync(kFull, v0, K - 1) : __shfl_sync(kFull, v1, K - 33);
  // a window whose last entry is within 2 eps of the K-th cannot exclude the entries outside it
  if (i_last != INT_MAX && !(a_last + 2.f * eps < a_k1)) {
    if (lane == 0 && p.n_second) atomicAdd(p.n_second, 1);
    row_topk_general(p, row, srow, qrow, ex, eps, lane);
    return;
  }

  // ---- which window entries can still be in the top K, and which of them S~ cannot order
  const bool rel0 = i0 != INT_MAX && (lane < K || !(v0 + 2.f * eps < a_k1));
  const bool rel1 = i1 != INT_MAX && (lane + 32 < K || !(v1 + 2.f * eps < a_k1));
  int c0 = 0, c1 = 0;
  bool need0 = rel0, need1 = rel1;
  if (!p.exact_values) {
    float pv0 = __shfl_up_sync(kFull, v0, 1);
    float pv1 = __shfl_up_sync(kFull, v1, 1);
    const float v0_last = __shfl_sync(kFull, v0, 31);
    if (lane == 0) pv1 = v0_last;
    const bool head0 = lane == 0 || (pv0 - v0 > 2.f * eps);
    const bool head1 = pv1 - v1 > 2.f * eps;
    const unsigned long long heads =
        (unsigned long long)__ballot_sync(kFull, head0) | ((unsigned long long)__ballot_sync(kFull, head1) << 32);
    const int p0 = lane, p1 = lane + 32;
    c0 = 63 - __clzll(heads & ((2ull << p0) - 1ull));
    c1 = 63 - __clzll(heads & (p1 == 63 ? ~0ull : ((2ull << p1) - 1ull)));
    const bool next_head0 = (heads >> (p0 + 1)) & 1ull;
    const bool next_head1 = p1 == 63 ? true : ((heads >> (p1 + 1)) & 1ull);
    need0 = rel0 && !(head0 && next_head0);  // singleton clusters keep their place without a chain
    need1 = rel1 && !(head1 && next_head1);
  }
  if (!rel0) c0 = INT_MAX;
  if (!rel1) c1 = INT_MAX;
  float e0, e1;
  warp_exact2(qrow, p.g, p.D, p.T, need0, i0 - p.col_base, need1, i1 - p.col_base, row_sm[wib], lane, e0, e1);
  float x0 = need0 ? e0 : v0, x1 = need1 ? e1 : v1;
  int j0 = i0, j1 = i1;

  // ---- 64-element bitonic sort on (cluster, score, index); position = lane + 32 * r
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {  // k == 64: register 0 against register 1 of the same lane, best first
        if (key_before(c1, x1, j1, c0, x0, j0)) {
          const int tc = c0; c0 = c1; c1 = tc;
          const float tv = x0; x0 = x1; x1 = tv;
          const int ti = j0; j0 = j1; j1 = ti;
        }
      } else {
        const bool lower = (lane & j) == 0;
        const bool up0 = k == 64 ? true : (k == 32 ? true : (lane & k) == 0);
        const bool up1 = k == 64 ? true : (k == 32 ? false : (lane & k) == 0);
        {
          const int tc = __shfl_xor_sync(kFull, c0, j);
          const float tv = __shfl_xor_sync(kFull, x0, j);
          const int ti = __shfl_xor_sync(kFull, j0, j);
          const bool mine_better = key_before(c0, x0, j0, tc, tv, ti);
          if (mine_better != (lower == up0)) { c0 = tc; x0 = tv; j0 = ti; }
        }
        {
          const int tc = __shfl_xor_sync(kFull, c1, j);
          const float tv = __shfl_xor_sync(kFull, x1, j);
          const int ti = __shfl_xor_sync(kFull, j1, j);
          const bool mine_better = key_before(c1, x1, j1, tc, tv, ti);
          if (mine_better != (lower == up1)) { c1 = tc; x1 = tv; j1 = ti; }
        }
      }
    }
  }
  if (c0 == INT_MAX) j0 = INT_MAX;
  if (c1 == INT_MAX) j1 = INT_MAX;
  write_topk(p, row, srow, x0, j0, x1, j1, lane);
}

__global__ void __launch_bounds__(kRowWarps * 32, 5)
row_rank_kernel(const RowParams p) {
  __shared__ float chain_sm[kRowWarps][kChainFloats];
  __shared__ int unsure_g[kRowWarps][352];  // < 64 carried (members + entries) + up to 256 per group
  const int wib = threadIdx.x >> 5;
  const int row = blockIdx.x * kRowWarps + wib;
  const int lane = threadIdx.x & 31;
  if (row >= p.qn) return;
  const float* srow = p.scores + size_t(row) * p.Gp;
  const float* qrow = p.q + size_t(row) * p.D;
  const int ex = p.exclude ? p.exclude[row] : -1;
  const int tg = p.target[row];
  float ts;
  if (p.tscore) {
    ts = p.tscore[row];
  } else {  // un-sharded call: the target row is here, its exact score is one more chain of this warp
    const int t = tg - p.col_base;
    ts = -INFINITY;
    if (t >= 0 && t < p.G) {  // warp-uniform
      float v = p.T == 1 ? warp_chains_few(qrow, p.g, p.D, lane == 0 ? t : -1, chain_sm[wib], lane)
                         : warp_chains(qrow, p.g, p.D, t * p.T + lane, chain_sm[wib], lane);
      if (p.T != 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
      }
      ts = __shfl_sync(kFull, v, 0);
    }
  }
  const float eps = row_eps(qrow, p.D, *p.gmax, lane);
  const int greg = (p.group && lane < p.M) ? p.group[size_t(row) * p.M + lane] : -2;
  int cnt = 0, gcnt = 0;
  unsigned n_unsure = 0;
  // entries within eps of the target score wait in a list (a stack) and are rescored, up to 32 at a time, with the
  // exact chain -- one call site, after each group of 256 entries.  Group members (the CIRR subset, validate.py:140-142)
  // are looked up directly, one per lane, instead of testing every gallery entry for membership; an unsure member goes
  // on the same list, tagged (~g).
  if (p.group) {
    const int ml = greg - p.col_base;
    bool valid = lane < p.M && ml >= 0 && ml < p.G && greg != ex && greg != tg;
    for (int mm = 0; mm < p.M; ++mm) {  // a member listed twice counts once
      const int o = __shfl_sync(kFull, greg, mm);
      if (mm < lane && o == greg) valid = false;
    }
    const float a = valid ? srow[ml] : -INFINITY;
    const bool before = valid && a > ts + eps;
    const bool unsure = valid && !before && !(a < ts - eps);
    gcnt = __popc(__ballot_sync(kFull, before));
    const unsigned um = __ballot_sync(kFull, unsure);
    if (unsure) unsure_g[wib][__popc(um & ((1u << lane) - 1u))] = ~ml;
    n_unsure = __popc(um);
  }
  for (int g0 = 0; g0 < p.G; g0 += 256) {
    float av[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * 32 + lane;
      av[u] = g < p.G ? srow[g] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * 32 + lane;
      const int idx = p.col_base + g;
      const bool ok = g < p.G && idx != ex && idx != tg;
      const float a = av[u];
      const bool before = ok && a > ts + eps;
      const bool unsure = ok && !before && !(a < ts - eps);
      const unsigned um = __ballot_sync(kFull, unsure);
      if (um) {
        if (unsure) unsure_g[wib][n_unsure + __popc(um & ((1u << lane) - 1u))] = g;
        n_unsure += __popc(um);
      }
      cnt += __popc(__ballot_sync(kFull, before));
    }
    const bool last = g0 + 256 >= p.G;
    while (n_unsure >= 32u || (last && n_unsure > 0u)) {
      const unsigned take = n_unsure < 32u ? n_unsure : 32u;
      __syncwarp();
      const bool have = unsigned(lane) < take;
      const int raw = have ? unsure_g[wib][n_unsure - take + lane] : 0;
      const bool member = raw < 0;
      const int g = member ? ~raw : raw;
      n_unsure -= take;
      const float e = warp_exact(qrow, p.g, p.D, p.T, have, g, chain_sm[wib], lane);
      const bool before = have && sorts_before(e, p.col_base + g, ts, tg);
      cnt += __popc(__ballot_sync(kFull, before && !member));
      gcnt += __popc(__ballot_sync(kFull, before && member));
    }
  }
  if (lane == 0) {
    if (cnt) atomicAdd(&p.rank[row], cnt);
    if (p.group_rank && gcnt) atomicAdd(&p.group_rank[row], gcnt);
  }
}

// ---- fused rank path (T = 1): rank_prep_kernel -> sim_tc_gemm_kernel<false, true> -> rank_resolve_kernel -------------
struct RankFusedParams {
  RowParams r;
  float* thr;     // [qn, 3]  decision band (ts - eps, ts + eps) and the exact target score ts
  int* cols;      // [qn, M1]
  float* col_s;   // [qn, M1]
  int* skip;      // [qn, 2]
  uint4* umask;   // [qn, g_tiles]   (1-CTA sweep)
  uint32_t* umask32;  // [qn, n_chunks] (pair sweep; non-NULL selects it)
  int M1, g_tiles, n_chunks;
};

// one warp per 32 queries: every lane owns one query and walks the exact fp32 chain of (q row, target row) -- the chain
// is sequential in k, so a warp per query would idle 31 lanes.  The warp loads 64-column chunks of its 32 + 32 rows with
// coalesced 16-byte loads (the next chunk is in flight in registers while the current one is walked from shared memory).
// Outputs per query: the decision band (ts - eps, ts + eps) and ts itself, the two columns left out of the count (the
// removed reference image, the target), the local columns of the subset members (validated exactly as row_rank_kernel
// does: in this shard, not the reference, not the target, first occurrence only).
constexpr int kPrepCols = 64;
constexpr int kPrepPitch = kPrepCols + 4;

__global__ void __launch_bounds__(32)
rank_prep_kernel(const RankFusedParams f) {
  __shared__ __align__(16) float sq[32][kPrepPitch];
  __shared__ __align__(16) float sg[32][kPrepPitch];
  const RowParams& p = f.r;
  const int lane = threadIdx.x;
  const int row0 = blockIdx.x * 32;
  const int row = row0 + lane;
  const bool live = row < p.qn;
  const int ex = (live && p.exclude) ? p.exclude[row] : -1;
  const int tg = live ? p.target[row] : -1;
  const int t = tg - p.col_base;
  const bool here = live && t >= 0 && t < p.G;
  // cooperative loads: step h covers rows 4 h + (lane >> 3), 16 floats (4 float4) per lane and row
  const int sub = lane >> 3, c16 = (lane & 7) * 8;  // this lane's 8 floats = 2 float4 inside a 64-column chunk
  int qrow_of[8], grow_of[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const int r = h * 4 + sub;
    qrow_of[h] = (row0 + r < p.qn) ? row0 + r : -1;
    const int tr = __shfl_sync(kFull, here ? t : -1, r);
    grow_of[h] = tr;
  }
  float4 qa[8][2], ga[8][2];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        qa[h][v] = qrow_of[h] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.q + size_t(qrow_of[h]) * p.D + k0 + c16) + v)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        ga[h][v] = grow_of[h] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.g + size_t(grow_of[h]) * p.D + k0 + c16) + v)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  float acc = 0.f, ss = 0.f;
  load_chunk(0);
  for (int k0 = 0; k0 < p.D; k0 += kPrepCols) {
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const int r = h * 4 + sub;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        *reinterpret_cast<float4*>(&sq[r][c16 + 4 * v]) = qa[h][v];
        *reinterpret_cast<float4*>(&sg[r][c16 + 4 * v]) = ga[h][v];
      }
    }
    __syncwarp();
    if (k0 + kPrepCols < p.D) load_chunk(k0 + kPrepCols);  // in flight while this chunk is walked
#pragma unroll 4
    for (int j = 0; j < kPrepCols; j += 4) {
      const float4 a = *reinterpret_cast<const float4*>(&sq[lane][j]);
      const float4 b = *reinterpret_cast<const float4*>(&sg[lane][j]);
      acc = fmaf(a.x, b.x, acc);
      acc = fmaf(a.y, b.y, acc);
      acc = fmaf(a.z, b.z, acc);
      acc = fmaf(a.w, b.w, acc);
      ss = fmaf(a.x, a.x, ss);
      ss = fmaf(a.y, a.y, ss);
      ss = fmaf(a.z, a.z, ss);
      ss = fmaf(a.w, a.w, ss);
    }
  }
  if (!live) return;
  const float ts = p.tscore ? p.tscore[row] : (here ? acc : -INFINITY);
  const float eps = kEpsRel * sqrtf(ss) * (*p.gmax) * 1.0001f;  // (another summation order than row_eps: keep a margin)
  f.thr[3 * row] = ts - eps;
  f.thr[3 * row + 1] = ts + eps;
  f.thr[3 * row + 2] = ts;
  const int exl = ex - p.col_base;
  f.skip[2 * row] = (ex >= 0 && exl >= 0 && exl < p.G) ? exl : -1;
  f.skip[2 * row + 1] = here ? t : -1;
  for (int mm = 0; mm < f.M1; ++mm) {
    const int gid = p.group[size_t(row) * p.M + mm];
    const int ml = gid - p.col_base;
    bool valid = ml >= 0 && ml < p.G && gid != ex && gid != tg;
    for (int o = 0; o < mm; ++o)  // a member listed twice counts once
      if (p.group[size_t(row) * p.M + o] == gid) valid = false;
    f.cols[size_t(row) * f.M1 + mm] = valid ? ml : -1;
    f.col_s[size_t(row) * f.M1 + mm] = -INFINITY;
  }
}

// the rare part of rank_resolve_kernel: rescore the listed undecided entries (up to 32 per round) with the exact chain
__device__ __noinline__ void resolve_batches(const RowParams& p, const float* qrow, int tg, bool last, int* list,
                                             float* chain, unsigned& n_unsure, float ts_exact, int& cnt, int& gcnt, int lane) {
  while (n_unsure >= 32u || (last && n_unsure > 0u)) {
    const unsigned take = n_unsure < 32u ? n_unsure : 32u;
    __syncwarp();
    const bool have = unsigned(lane) < take;
    const int raw = have ? list[n_unsure - take + lane] : 0;
    const bool member = raw < 0;
    const int g = member ? ~raw : raw;
    n_unsure -= take;
    const float e = warp_exact(qrow, p.g, p.D, p.T, have, g, chain, lane);
    const bool before = have && sorts_before(e, p.col_base + g, ts_exact, tg);
    cnt += __popc(__ballot_sync(kFull, before && !member));
    gcnt += __popc(__ballot_sync(kFull, before && member));
  }
}

// one warp per query: rescore the undecided entries recorded by the sweep's epilogue (exact fp32 chains, up to 32 at a
// time) and decide the subset members from their extracted S~ (exact chain only when within eps)
__global__ void __launch_bounds__(kRowWarps * 32, 5)
rank_resolve_kernel(const RankFusedParams f) {
  __shared__ float chain_sm[kRowWarps][kChainFloats];
  __shared__ int unsure_g[kRowWarps][352];
  const RowParams& p = f.r;
  const int wib = threadIdx.x >> 5;
  const int row = blockIdx.x * kRowWarps + wib;
  const int lane = threadIdx.x & 31;
  if (row >= p.qn) return;
  const float* qrow = p.q + size_t(row) * p.D;
  const int tg = p.target[row];
  const float lo = f.thr[3 * row], hi = f.thr[3 * row + 1];
  const float ts_exact = f.thr[3 * row + 2];  // the exact fp32 chain of the target (rank_prep_kernel) or the caller's value
  int cnt = 0, gcnt = 0;
  unsigned n_unsure = 0;
  if (f.M1 > 0) {
    const int ml = lane < f.M1 ? f.cols[size_t(row) * f.M1 + lane] : -1;
    const bool valid = ml >= 0;
    const float a = valid ? f.col_s[size_t(row) * f.M1 + lane] : -INFINITY;
    const bool before = valid && a > hi;
    const bool unsure = valid && !before && !(a < lo);
    gcnt = __popc(__ballot_sync(kFull, before));
    const unsigned um = __ballot_sync(kFull, unsure);
    if (unsure) unsure_g[wib][__popc(um & ((1u << lane) - 1u))] = ~ml;
    n_unsure = __popc(um);
  }
  auto flush = [&](bool last) {
    if (n_unsure >= 32u || (last && n_unsure > 0u))  // rare: kept out of line so the common path stays in registers
      resolve_batches(p, qrow, tg, last, unsure_g[wib], chain_sm[wib], n_unsure, ts_exact, cnt, gcnt, lane);
  };
  if (f.umask32) {
    // pair sweep: one 32-bit mask per 32-column chunk; lane == chunk, chunks with a non-zero mask are visited one at a time
    for (int c0 = 0; c0 < f.n_chunks; c0 += 32) {
      const int c = c0 + lane;
      const uint32_t w = c < f.n_chunks ? f.umask32[size_t(row) * f.n_chunks + c] : 0u;
      unsigned nz = __ballot_sync(kFull, w != 0u);
      while (nz) {
        const int src = __ffs(nz) - 1;
        nz &= nz - 1;
        const uint32_t word = __shfl_sync(kFull, w, src);
        const bool unsure = (word >> lane) & 1u;
        const unsigned um = __ballot_sync(kFull, unsure);
        if (unsure) unsure_g[wib][n_unsure + __popc(um & ((1u << lane) - 1u))] = (c0 + src) * 32 + lane;
        n_unsure += __popc(um);
        flush(false);  // keeps the list below 32 + 32 entries (+ the members)
      }
    }
  } else {
  // each lane owns 4 of the 128 bits of a tile mask; tiles are visited one at a time only when their mask is non-zero
  for (int t0 = 0; t0 < f.g_tiles; t0 += 32) {
    const int t = t0 + lane;
    uint4 m = make_uint4(0u, 0u, 0u, 0u);
    if (t < f.g_tiles) m = f.umask[size_t(row) * f.g_tiles + t];
    unsigned nz = __ballot_sync(kFull, (m.x | m.y | m.z | m.w) != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      const uint32_t w0 = __shfl_sync(kFull, m.x, src), w1 = __shfl_sync(kFull, m.y, src);
      const uint32_t w2 = __shfl_sync(kFull, m.z, src), w3 = __shfl_sync(kFull, m.w, src);
      const uint32_t word = (lane >> 3) == 0 ? w0 : (lane >> 3) == 1 ? w1 : (lane >> 3) == 2 ? w2 : w3;
      const uint32_t nib = (word >> ((lane & 7) * 4)) & 0xFu;  // bits 4 lane .. 4 lane + 3 of the 128
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const bool unsure = (nib >> b) & 1u;
        const unsigned um = __ballot_sync(kFull, unsure);
        if (um) {
          if (unsure) unsure_g[wib][n_unsure + __popc(um & ((1u << lane) - 1u))] = (t0 + src) * kBN + lane * 4 + b;
          n_unsure += __popc(um);
        }
      }
      flush(false);  // keeps the list below 32 + 128 entries (+ the members)
    }
  }
  }
  flush(true);
  if (lane == 0) {
    if (cnt) atomicAdd(&p.rank[row], cnt);
    if (p.group_rank && gcnt) atomicAdd(&p.group_rank[row], gcnt);
  }
}

// score[b] of the target entry with the exact chain (same bits as target_score_kernel of sim_rank.cu), one warp per query
__global__ void __launch_bounds__(kRowWarps * 32)
target_score_warp_kernel(const float* __restrict__ q, const float* __restrict__ g, int Q, int G, int D, int T,
                         const int* __restrict__ target, int col_base, float* __restrict__ score) {
  __shared__ float chain_sm[kRowWarps][kChainFloats];
  const int wib = threadIdx.x >> 5;
  const int row = blockIdx.x * kRowWarps + wib;
  const int lane = threadIdx.x & 31;
  if (row >= Q) return;
  const int t = target[row] - col_base;
  float v = -INFINITY;
  if (t >= 0 && t < G) {  // warp-uniform
    v = T == 1 ? warp_chains_few(q + size_t(row) * D, g, D, lane == 0 ? t : -1, chain_sm[wib], lane)
               : warp_chains(q + size_t(row) * D, g, D, t * T + lane, chain_sm[wib], lane);
    if (T != 1) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
    }
  }
  if (lane == 0) score[row] = v;
}

// ------------------------------------------------------------------------------------------------ host side
struct TcLayout {
  size_t qh, ql, gh, gl, scores, misc, total;
  int qb;  // queries per block (multiple of 128)
  int Gp;
};

TcLayout tc_layout(int Q, int G, int D, int T) {
  TcLayout L{};
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  L.Gp = (G + 3) / 4 * 4;
  size_t budget = size_t(512) << 20;  // score block budget (SPN_TC_SCORE_BLOCK_MB overrides: tests force several blocks)
  if (const char* e = getenv("SPN_TC_SCORE_BLOCK_MB")) {
    const long v = atol(e);
    if (v >= 1 && v <= 65536) budget = size_t(v) << 20;
  }
  long long qb = (long long)(budget / (size_t(L.Gp) * 4)) / 128 * 128;
  if (qb < 128) qb = 128;
  const long long qpad = (long long)(Q + 127) / 128 * 128;
  if (qb > qpad) qb = qpad;
  L.qb = int(qb);
  size_t o = 0;
  L.qh = o; o += up(size_t(Q) * D * 2);
  L.ql = o; o += up(size_t(Q) * D * 2);
  L.gh = o; o += up(size_t(G) * T * D * 2);
  L.gl = o; o += up(size_t(G) * T * D * 2);
  L.scores = o; o += up(size_t(L.qb) * L.Gp * 4);
  L.misc = o; o += 256;
  L.total = o;
  return L;
}

int tc_check(const float* q, const float* g, int Q, int G, int D, int T) {
  SPN_REQUIRE(Q > 0 && G > 0, SPN_ERR_BAD_SHAPE, "retrieval needs Q > 0 and G > 0 (Q=%d G=%d)", Q, G);
  SPN_REQUIRE(D >= 64 && D % 64 == 0, SPN_ERR_BAD_SHAPE, "tensor-core retrieval needs D %% 64 == 0 (D=%d)", D);
  SPN_REQUIRE(T == 1 || T == 32, SPN_ERR_BAD_SHAPE, "tokens per entry T=%d unsupported (1 or 32)", T);
  SPN_REQUIRE((long long)G * T < (1ll << 31), SPN_ERR_BAD_SHAPE, "gallery too large: G*T=%lld", (long long)G * T);
  SPN_REQUIRE(aligned16(q) && aligned16(g), SPN_ERR_BAD_ALIGN, "retrieval operands must be 16-byte aligned");
  return SPN_OK;
}

size_t tc_gemm_smem() { return 1024 + size_t(kTcStages) * kStageBytes + 8 * kSlotBytes + 256; }

// split both operands, then per query block: tensor-core sweep + row kernel (`rank_mode` selects which)
int tc_run(bool rank_mode, const float* q, const float* g, int Q, int G, int D, int T, RowParams rp, int* n_second_out,
           void* ws, size_t ws_bytes, cudaStream_t stream) {
  int rc = tc_check(q, g, Q, G, D, T);
  if (rc != SPN_OK) return rc;
  const TcLayout L = tc_layout(Q, G, D, T);
  SPN_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 255u) == 0 && ws_bytes >= L.total, SPN_ERR_WORKSPACE,
              "tensor-core retrieval workspace too small or misaligned: %zu < %zu", ws_bytes, L.total);
  uint8_t* w = static_cast<uint8_t*>(ws);
  auto* qh = reinterpret_cast<__nv_bfloat16*>(w + L.qh);
  auto* ql = reinterpret_cast<__nv_bfloat16*>(w + L.ql);
  auto* gh = reinterpret_cast<__nv_bfloat16*>(w + L.gh);
  auto* gl = reinterpret_cast<__nv_bfloat16*>(w + L.gl);
  float* scores = reinterpret_cast<float*>(w + L.scores);
  float* gmax = reinterpret_cast<float*>(w + L.misc);
  int* n_second = reinterpret_cast<int*>(w + L.misc + 16);
  SPN_CUDA_CHECK(cudaMemsetAsync(w + L.misc, 0, 256, stream));
  const long long grows = (long long)G * T;
  split_rows_kernel<<<unsigned((Q + 7) / 8), 256, 0, stream>>>(q, Q, D, qh, ql, nullptr);
  split_rows_kernel<<<unsigned((grows + 7) / 8), 256, 0, stream>>>(g, grows, D, gh, gl, gmax);
  SPN_CUDA_CHECK(cudaGetLastError());

  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  // pair sweep (default): cluster tile = 256 queries x tile_n gallery rows; tile_n (multiple of 32, <= 256) minimises
  // rounds x instruction time, an SS M256 x N x K16 instruction costing ~N / 2 + 43 cycles (profiles/r01_exp_mma_rate_ss_ts)
  bool pair = true;
  if (const char* e = getenv("SPN_TC_SWEEP_CG")) {
    if (atoi(e) == 1) pair = false;
  }
  const int qb_for_plan = Q < L.qb ? Q : L.qb;
  const int q_tiles2_plan = (qb_for_plan + 2 * kBM - 1) / (2 * kBM);
  const int n_clusters = sms / 2;
  int tile_n = 256;
  {
    double best = 1e30;
    for (int n = 256; n >= 128; n -= 32) {
      const long long tiles = (long long)q_tiles2_plan * ((grows + n - 1) / n);
      const long long rounds = (tiles + n_clusters - 1) / n_clusters;
      const double cost = double(rounds) * (n / 2 + 43);
      if (cost < best - 1e-9) {
        best = cost;
        tile_n = n;
      }
    }
  }
  if (const char* e = getenv("SPN_TC_TILE_N")) {
    const int v = atoi(e);
    if (v >= 32 && v <= 256 && v % 32 == 0) tile_n = v;
  }

  CUtensorMap tqh, tql, tgh, tgl, ts;
  if ((rc = make_tmap_bf16_2d(&tqh, qh, uint64_t(Q), uint64_t(D), kBM)) != SPN_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tql, ql, uint64_t(Q), uint64_t(D), kBM)) != SPN_OK) return rc;
  const uint32_t g_box = pair ? uint32_t(tile_n / 2) : uint32_t(kBN);  // pair sweep: each CTA stages half of the tile's rows
  if ((rc = make_tmap_bf16_2d(&tgh, gh, uint64_t(grows), uint64_t(D), g_box)) != SPN_OK) return rc;
  if ((rc = make_tmap_bf16_2d(&tgl, gl, uint64_t(grows), uint64_t(D), g_box)) != SPN_OK) return rc;

  const size_t smem = tc_gemm_smem();
  // per launch, like the bank-CE kernels: the attribute is per device and the call costs a few microseconds of host time
  SPN_REQUIRE(size_t(device_max_smem_optin()) >= smem, SPN_ERR_CUDA, "device offers %d B of shared memory, %zu needed",
              device_max_smem_optin(), smem);
  if (pair) {
    SPN_CUDA_CHECK(cudaFuncSetAttribute(sim_tc_gemm2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    SPN_CUDA_CHECK(cudaFuncSetAttribute(sim_tc_gemm2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    SPN_CUDA_CHECK(cudaFuncSetAttribute(sim_tc_gemm2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  } else if (T == 1) {
    SPN_CUDA_CHECK(cudaFuncSetAttribute(sim_tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  } else {
    SPN_CUDA_CHECK(cudaFuncSetAttribute(sim_tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  }
  auto launch_pair_sweep = [&](int variant, const TcGemmParams& gp2) -> int {  // 0: S~ store, 1: token max, 2: rank count
    const long long tiles = (long long)gp2.q_tiles2 * gp2.g_tiles2;
    const int clusters = int(tiles < n_clusters ? tiles : n_clusters);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(2 * clusters), 1, 1);
    cfg.blockDim = dim3(kTcThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (variant == 0) SPN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sim_tc_gemm2_kernel<false, false>, tqh, tql, tgh, tgl, ts, gp2));
    else if (variant == 1) SPN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sim_tc_gemm2_kernel<true, false>, tqh, tql, tgh, tgl, ts, gp2));
    else SPN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sim_tc_gemm2_kernel<false, true>, tqh, tql, tgh, tgl, ts, gp2));
    return SPN_OK;
  };

  // rank, T = 1: the fused path -- no S~ matrix; the sweep's epilogue counts and the resolve kernel rescored the undecided
  // entries (SPN_TC_RANK_FUSED=0 keeps the S~ + row_rank_kernel path, also taken when the workspace block is too small)
  bool rank_fused = rank_mode && T == 1;
  if (const char* e = getenv("SPN_TC_RANK_FUSED")) {
    if (atoi(e) == 0) rank_fused = false;
  }
  const int M1 = rp.group ? rp.M : 0;
  const int g_tiles_all = int((grows + kBN - 1) / kBN);
  const int n_chunks = int((grows + 31) / 32);
  auto up256 = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t fused_need = up256(size_t(L.qb) * 12) + 2 * up256(size_t(L.qb) * (M1 > 0 ? M1 : 1) * 4) + up256(size_t(L.qb) * 8) +
                            up256(size_t(L.qb) * g_tiles_all * 16);  // (the chunk-mask layout of the pair sweep is no larger)
  if (fused_need > up256(size_t(L.qb) * L.Gp * 4)) rank_fused = false;
  if (rank_fused && !pair)
    SPN_CUDA_CHECK(cudaFuncSetAttribute(sim_tc_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));

  for (int qb0 = 0; qb0 < Q; qb0 += L.qb) {
    const int qn = (Q - qb0 < L.qb) ? Q - qb0 : L.qb;
    if ((rc = make_tmap_f32_2d(&ts, scores, uint64_t(qn), uint64_t(L.Gp), 32)) != SPN_OK) return rc;
    TcGemmParams gp{};
    gp.q_row0 = qb0; gp.qn = qn; gp.G = G; gp.Gp = L.Gp; gp.kblocks = D / kBK;
    gp.q_tiles = (qn + kBM - 1) / kBM;
    gp.g_tiles = int((grows + kBN - 1) / kBN);
    gp.nterms = 3;
    gp.scores = scores;
    gp.tile_n = tile_n;
    gp.q_tiles2 = (qn + 2 * kBM - 1) / (2 * kBM);
    gp.g_tiles2 = int((grows + tile_n - 1) / tile_n);
    gp.n_chunks = n_chunks;
    const long long n_tiles = (long long)gp.q_tiles * gp.g_tiles;
    const int grid = int(n_tiles < sms ? n_tiles : sms);
    RowParams r = rp;
    r.q = q + size_t(qb0) * D; r.g = g; r.scores = scores; r.gmax = gmax; r.qn = qn; r.G = G; r.Gp = L.Gp; r.D = D; r.T = T;
    r.n_second = n_second;
    if (r.exclude) r.exclude += qb0;
    if (rank_mode) {
      r.target += qb0;
      if (r.tscore) r.tscore += qb0;
      if (r.group) r.group += size_t(qb0) * r.M;
      r.rank += qb0;
      if (r.group_rank) r.group_rank += qb0;
    }
    if (rank_fused) {
      RankFusedParams f{};
      uint8_t* o = reinterpret_cast<uint8_t*>(scores);
      f.thr = reinterpret_cast<float*>(o); o += up256(size_t(L.qb) * 12);
      f.cols = reinterpret_cast<int*>(o); o += up256(size_t(L.qb) * (M1 > 0 ? M1 : 1) * 4);
      f.col_s = reinterpret_cast<float*>(o); o += up256(size_t(L.qb) * (M1 > 0 ? M1 : 1) * 4);
      f.skip = reinterpret_cast<int*>(o); o += up256(size_t(L.qb) * 8);
      f.umask = reinterpret_cast<uint4*>(o);
      f.M1 = M1; f.g_tiles = gp.g_tiles; f.r = r;
      const int rblocks = (qn + kRowWarps - 1) / kRowWarps;
      rank_prep_kernel<<<(qn + 31) / 32, 32, 0, stream>>>(f);
      f.umask32 = pair ? reinterpret_cast<uint32_t*>(f.umask) : nullptr;
      f.n_chunks = n_chunks;
      gp.thr = f.thr; gp.cols = f.cols; gp.col_s = f.col_s; gp.M1 = M1; gp.skip = f.skip; gp.umask = f.umask; gp.rank = r.rank;
      gp.umask32 = f.umask32;
      if (pair) {
        if ((rc = launch_pair_sweep(2, gp)) != SPN_OK) return rc;
      } else {
        sim_tc_gemm_kernel<false, true><<<grid, kTcThreads, smem, stream>>>(tqh, tql, tgh, tgl, ts, gp);
      }
      rank_resolve_kernel<<<rblocks, kRowWarps * 32, 0, stream>>>(f);
      SPN_CUDA_CHECK(cudaGetLastError());
      continue;
    }
    if (pair) {
      if ((rc = launch_pair_sweep(T == 1 ? 0 : 1, gp)) != SPN_OK) return rc;
    } else if (T == 1) {
      sim_tc_gemm_kernel<false><<<grid, kTcThreads, smem, stream>>>(tqh, tql, tgh, tgl, ts, gp);
    } else {
      sim_tc_gemm_kernel<true><<<grid, kTcThreads, smem, stream>>>(tqh, tql, tgh, tgl, ts, gp);
    }
    SPN_CUDA_CHECK(cudaGetLastError());

    if (rank_mode) {
      row_rank_kernel<<<(qn + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, stream>>>(r);
    } else {
      r.out_val += size_t(qb0) * r.K; r.out_idx += size_t(qb0) * r.K;
      row_topk_kernel<<<(qn + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, stream>>>(r);
    }
    SPN_CUDA_CHECK(cudaGetLastError());
  }
  if (n_second_out)
    SPN_CUDA_CHECK(cudaMemcpyAsync(n_second_out, n_second, sizeof(int), cudaMemcpyDeviceToDevice, stream));
  return SPN_OK;
}

}  // namespace

size_t sim_tc_ws_bytes(int Q, int G, int D, int T) { return tc_layout(Q, G, D, T).total; }

int sim_tc_debug_layout(int Q, int G, int D, int T, size_t* scores_offset, int* pitch, int* block_rows) {
  const TcLayout L = tc_layout(Q, G, D, T);
  *scores_offset = L.scores;
  *pitch = L.Gp;
  *block_rows = L.qb;
  return SPN_OK;
}

int sim_topk_tc(const float* q, const float* g, int Q, int G, int D, int T, int K, const int* exclude, int col_base,
                float* out_val, int* out_idx, int exact_values, int* n_second, void* ws, size_t ws_bytes,
                cudaStream_t stream) {
  SPN_REQUIRE(K >= 1 && K <= 56, SPN_ERR_BAD_SHAPE, "tensor-core top-K supports 1 <= K <= 56 (K=%d): 64 candidates are rescored", K);
  RowParams rp{};
  rp.K = K; rp.col_base = col_base; rp.exclude = exclude; rp.out_val = out_val; rp.out_idx = out_idx;
  rp.exact_values = exact_values;
  return tc_run(false, q, g, Q, G, D, T, rp, n_second, ws, ws_bytes, stream);
}

int sim_rank_tc(const float* q, const float* g, int Q, int G, int D, int T, const int* target, const float* tscore,
                const int* exclude, const int* group, int M, int col_base, int* rank, int* group_rank, void* ws,
                size_t ws_bytes, cudaStream_t stream) {
  SPN_REQUIRE(target && rank, SPN_ERR_BAD_ARG, "sim_rank_tc needs target and rank");
  SPN_REQUIRE(!group || (M >= 1 && M <= 32 && group_rank), SPN_ERR_BAD_SHAPE, "sim_rank: group needs 1 <= M <= 32 and group_rank");
  RowParams rp{};
  rp.K = 1; rp.col_base = col_base; rp.exclude = exclude; rp.target = target; rp.tscore = tscore; rp.group = group; rp.M = M;
  rp.rank = rank; rp.group_rank = group_rank;
  return tc_run(true, q, g, Q, G, D, T, rp, nullptr, ws, ws_bytes, stream);
}

int sim_target_score_warp(const float* q, const float* g, int Q, int G, int D, int T, const int* target, int col_base,
                          float* score, cudaStream_t stream) {
  SPN_REQUIRE(Q > 0 && G > 0 && D >= 64 && D % 64 == 0 && (T == 1 || T == 32), SPN_ERR_BAD_SHAPE,
              "sim_target_score_warp: bad shape Q=%d G=%d D=%d T=%d", Q, G, D, T);
  target_score_warp_kernel<<<(Q + kRowWarps - 1) / kRowWarps, kRowWarps * 32, 0, stream>>>(q, g, Q, G, D, T, target, col_base,
                                                                                       score);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
