// Retrieval scoring: fused fp32 similarity + per-row top-K / rank-of-target.
// Replaces `distances = 1 - q @ g.T; torch.argsort(distances)` (clip4cir/validate.py:126-127 and the
// 15 sibling call sites) and the max-over-32-tokens scoring of Blip2QformerCirAlignPrompt.inference
// (...align_prompt.py:311-319).  The [Q, G] score matrix only exists as 64 x 64 tiles in shared memory.
//
// Arithmetic is exact fp32 FMA in ascending-k order (one accumulator per score), so a score is
// bit-identical wherever it is computed (tile kernel or spn_sim_target_score) and ranks are a
// deterministic function of the inputs: descending score, ties -> lower global index first.
//
// Tile kernel: 256 threads, 64 queries x 64 gallery rows x 16-deep k-slices, 4x4 register micro-tile;
// epilogue: warp w owns rows 8w..8w+7; the running top-K list of a row lives in registers spread over
// the 32 lanes (positions lane and lane+32) and is updated with ballot/shuffle insertion.
#include <limits.h>
#include <math.h>

#include "bank_ce.cuh"
#include "sim_rank.cuh"
#include "topk_list.cuh"

namespace spn {
namespace {

constexpr int TQ = 64, TG = 64, TKK = 16;
constexpr int kSimThreads = 256;
struct SimParams {
  const float* q;
  const float* g;
  int Q, G, D, T, K;
  int col_base;
  int tiles_per_chunk;  // gallery tiles (64 rows) per blockIdx.y
  const int* exclude;
  // top-k cursor (nullable): only candidates that sort strictly AFTER (after_val[q], after_idx[q]) are eligible --
  // page p + 1 of a deep top-K starts where page p ended
  const float* after_val;
  const int* after_idx;
  // top-k outputs [chunks][Q][K]
  float* out_val;
  int* out_idx;
  // rank
  const int* target;
  const float* tscore;
  const int* group;
  int M;
  int* rank;
  int* group_rank;
};

// MODE 0: top-K, MODE 1: rank-of-target, MODE 2: dense score matrix (compat).  T1: true when T == 1 (else 32 token rows per entry).
template <int MODE, bool T1>
__global__ void __launch_bounds__(kSimThreads)
sim_tile_kernel(const SimParams p) {
  __shared__ __align__(16) float As[TKK][TQ + 4];
  __shared__ __align__(16) float Bs[TKK][TG + 4];
  __shared__ float S[TQ][TG + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * TQ;
  const long long grows = (long long)p.G * p.T;
  const int n_tiles = int((grows + TG - 1) / TG);
  const int tile_begin = blockIdx.y * p.tiles_per_chunk;
  const int tile_end = min(n_tiles, tile_begin + p.tiles_per_chunk);

  // per-warp-row state (row = warp*8 + r)
  float lv0[8], lv1[8];
  int li0[8], li1[8];
  int cnt[8], gcnt[8], greg[8], tg[8], ex[8], ai[8];
  float ts[8], av[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    lv0[r] = lv1[r] = -INFINITY;
    li0[r] = li1[r] = INT_MAX;
    cnt[r] = gcnt[r] = 0;
    const int qq = q0 + warp * 8 + r;
    const bool ok = qq < p.Q;
    ex[r] = (ok && p.exclude) ? p.exclude[qq] : -1;
    av[r] = (MODE == 0 && ok && p.after_val) ? p.after_val[qq] : INFINITY;  // +inf / -1 sorts before everything
    ai[r] = (MODE == 0 && ok && p.after_val) ? p.after_idx[qq] : -1;
    if (MODE == 1) {
      tg[r] = ok ? p.target[qq] : -1;
      ts[r] = ok ? p.tscore[qq] : INFINITY;
      greg[r] = (ok && p.group && lane < p.M) ? p.group[size_t(qq) * p.M + lane] : -2;
    }
  }

  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const long long g0 = (long long)tile * TG;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < p.D; k0 += TKK) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (q0 + lrow < p.Q) a = *reinterpret_cast<const float4*>(p.q + size_t(q0 + lrow) * p.D + k0 + lk);
      if (g0 + lrow < grows) b = *reinterpret_cast<const float4*>(p.g + size_t(g0 + lrow) * p.D + k0 + lk);
      __syncthreads();  // previous slice fully consumed
      As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
      Bs[lk + 0][lrow] = b.x; Bs[lk + 1][lrow] = b.y; Bs[lk + 2][lrow] = b.z; Bs[lk + 3][lrow] = b.w;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TKK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w};
        const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) S[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();

    // ---- epilogue on the 64 x 64 score tile
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = warp * 8 + r;
      if (q0 + row >= p.Q) continue;  // warp-uniform
      float c0v, c1v;
      int c0i, c1i;
      bool c0ok, c1ok;
      if (T1) {
        c0v = S[row][lane];
        c1v = S[row][lane + 32];
        c0i = p.col_base + int(g0) + lane;
        c1i = c0i + 32;
        c0ok = g0 + lane < grows;
        c1ok = g0 + lane + 32 < grows;
      } else {
        float m0 = S[row][lane], m1 = S[row][lane + 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          m0 = fmaxf(m0, __shfl_xor_sync(kFull, m0, o));
          m1 = fmaxf(m1, __shfl_xor_sync(kFull, m1, o));
        }
        const int e0 = int(g0 / 32);
        c0v = m0; c1v = m1;
        c0i = p.col_base + e0; c1i = c0i + 1;
        c0ok = lane == 0 && e0 < p.G;
        c1ok = lane == 0 && e0 + 1 < p.G;
      }
      c0ok = c0ok && c0i != ex[r];
      c1ok = c1ok && c1i != ex[r];
      if (MODE == 2) {  // dense [Q, G] output, out_val row pitch = G (exclude is ignored)
        float* orow = p.out_val + size_t(q0 + row) * p.G - p.col_base;
        if (T1) {
          if (g0 + lane < grows) orow[c0i] = c0v;
          if (g0 + lane + 32 < grows) orow[c1i] = c1v;
        } else if (lane == 0) {
          if (c0i - p.col_base < p.G) orow[c0i] = c0v;
          if (c1i - p.col_base < p.G) orow[c1i] = c1v;
        }
      } else if (MODE == 0) {
        c0ok = c0ok && sorts_before(av[r], ai[r], c0v, c0i);
        c1ok = c1ok && sorts_before(av[r], ai[r], c1v, c1i);
        topk_offer(c0v, c0i, c0ok, p.K, lv0[r], li0[r], lv1[r], li1[r], lane);
        topk_offer(c1v, c1i, c1ok, p.K, lv0[r], li0[r], lv1[r], li1[r], lane);
      } else {
        const bool b0 = c0ok && c0i != tg[r] && sorts_before(c0v, c0i, ts[r], tg[r]);
        const bool b1 = c1ok && c1i != tg[r] && sorts_before(c1v, c1i, ts[r], tg[r]);
        cnt[r] += __popc(__ballot_sync(kFull, b0)) + __popc(__ballot_sync(kFull, b1));
        if (p.group) {
          bool in0 = false, in1 = false;
          for (int mm = 0; mm < p.M; ++mm) {
            const int gm = __shfl_sync(kFull, greg[r], mm);
            in0 |= (gm == c0i);
            in1 |= (gm == c1i);
          }
          gcnt[r] += __popc(__ballot_sync(kFull, b0 && in0)) + __popc(__ballot_sync(kFull, b1 && in1));
        }
      }
    }
    // S is rewritten only after the next tile's k-loop, which starts with __syncthreads()
  }

#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int qq = q0 + warp * 8 + r;
    if (qq >= p.Q) continue;
    if (MODE == 0) {
      const size_t o = (size_t(blockIdx.y) * p.Q + qq) * p.K;
      if (lane < p.K) {
        const bool real = li0[r] != INT_MAX;
        p.out_val[o + lane] = real ? lv0[r] : -INFINITY;
        p.out_idx[o + lane] = real ? li0[r] : -1;
      }
      if (lane + 32 < p.K) {
        const bool real = li1[r] != INT_MAX;
        p.out_val[o + lane + 32] = real ? lv1[r] : -INFINITY;
        p.out_idx[o + lane + 32] = real ? li1[r] : -1;
      }
    } else if (MODE == 1 && lane == 0) {
      if (cnt[r]) atomicAdd(&p.rank[qq], cnt[r]);
      if (p.group_rank && gcnt[r]) atomicAdd(&p.group_rank[qq], gcnt[r]);
    }
  }
}

// one warp per query: merge R lists of K
__global__ void merge_topk_kernel(const float* __restrict__ vals, const int* __restrict__ idx, int R, int Q, int K,
                                  float* __restrict__ out_val, int* __restrict__ out_idx) {
  const int qq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (qq >= Q) return;
  float v0 = -INFINITY, v1 = -INFINITY;
  int i0 = INT_MAX, i1 = INT_MAX;
  for (int r = 0; r < R; ++r) {
    const size_t o = (size_t(r) * Q + qq) * K;
    for (int pbase = 0; pbase < K; pbase += 32) {
      const int pp = pbase + lane;
      const bool in = pp < K;
      const float cv = in ? vals[o + pp] : -INFINITY;
      const int ci = in ? idx[o + pp] : -1;
      topk_offer(cv, ci, in && ci >= 0, K, v0, i0, v1, i1, lane);
    }
  }
  const size_t o = size_t(qq) * K;
  if (lane < K) {
    const bool real = i0 != INT_MAX;
    out_val[o + lane] = real ? v0 : -INFINITY;
    out_idx[o + lane] = real ? i0 : -1;
  }
  if (lane + 32 < K) {
    const bool real = i1 != INT_MAX;
    out_val[o + lane + 32] = real ? v1 : -INFINITY;
    out_idx[o + lane + 32] = real ? i1 : -1;
  }
}

// one thread per query; same ascending-k FMA chain as the tile kernel -> bit-identical scores
__global__ void target_score_kernel(const float* __restrict__ q, const float* __restrict__ g, int Q, int G, int D,
                                    int T, const int* __restrict__ target, int col_base, float* __restrict__ score) {
  const int qq = blockIdx.x * blockDim.x + threadIdx.x;
  if (qq >= Q) return;
  const int t = target[qq] - col_base;
  if (t < 0 || t >= G) {
    score[qq] = -INFINITY;
    return;
  }
  const float* a = q + size_t(qq) * D;
  float best = -INFINITY;
  for (int tok = 0; tok < T; ++tok) {
    const float* b = g + (size_t(t) * T + tok) * D;
    float acc = 0.f;
    for (int k = 0; k < D; ++k) acc = fmaf(a[k], b[k], acc);
    best = fmaxf(best, acc);
  }
  score[qq] = best;
}

int plan_chunks(int Q, long long grows, int* chunks, int* tiles_per_chunk) {
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int q_tiles = (Q + TQ - 1) / TQ;
  const int n_tiles = int((grows + TG - 1) / TG);
  int c = (2 * sms + q_tiles - 1) / q_tiles;  // aim at ~2 CTAs per SM
  if (c > n_tiles) c = n_tiles;
  if (c > 64) c = 64;
  if (c < 1) c = 1;
  const int tpc = (n_tiles + c - 1) / c;
  *tiles_per_chunk = tpc;
  *chunks = (n_tiles + tpc - 1) / tpc;
  return q_tiles;
}

int check_common(const float* q, const float* g, int Q, int G, int D, int T) {
  SPN_REQUIRE(Q > 0 && G > 0, SPN_ERR_BAD_SHAPE, "retrieval needs Q > 0 and G > 0 (Q=%d G=%d)", Q, G);
  SPN_REQUIRE(D > 0 && D % 16 == 0, SPN_ERR_BAD_SHAPE, "retrieval needs D %% 16 == 0 (D=%d)", D);
  SPN_REQUIRE(T == 1 || T == 32, SPN_ERR_BAD_SHAPE, "tokens per entry T=%d unsupported (1 or 32)", T);
  SPN_REQUIRE(aligned16(q) && aligned16(g), SPN_ERR_BAD_ALIGN, "retrieval operands must be 16-byte aligned");
  return SPN_OK;
}

}  // namespace

size_t sim_topk_ws_bytes(int Q, int G, int T, int K) {
  int chunks, tpc;
  plan_chunks(Q, (long long)G * T, &chunks, &tpc);
  return ((size_t(chunks) * Q * K * 8 + 255) / 256) * 256;
}

int sim_topk(const float* q, const float* g, int Q, int G, int D, int T, int K, const int* exclude, int col_base,
             float* out_val, int* out_idx, void* ws, size_t ws_bytes, cudaStream_t stream, const float* after_val,
             const int* after_idx) {
  SPN_REQUIRE((after_val == nullptr) == (after_idx == nullptr), SPN_ERR_BAD_ARG, "top-K cursor needs both after_val and after_idx");
  int rc = check_common(q, g, Q, G, D, T);
  if (rc != SPN_OK) return rc;
  SPN_REQUIRE(K >= 1 && K <= 64, SPN_ERR_BAD_SHAPE, "top-K supports 1 <= K <= 64 (K=%d)", K);
  int chunks, tpc;
  const int q_tiles = plan_chunks(Q, (long long)G * T, &chunks, &tpc);
  const size_t need = ((size_t(chunks) * Q * K * 8 + 255) / 256) * 256;
  SPN_REQUIRE(chunks == 1 || (ws && ws_bytes >= need), SPN_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, need);
  SimParams p{};
  p.q = q; p.g = g; p.Q = Q; p.G = G; p.D = D; p.T = T; p.K = K; p.col_base = col_base;
  p.tiles_per_chunk = tpc; p.exclude = exclude; p.after_val = after_val; p.after_idx = after_idx;
  if (chunks == 1) {
    p.out_val = out_val; p.out_idx = out_idx;
  } else {
    p.out_val = static_cast<float*>(ws);
    p.out_idx = reinterpret_cast<int*>(p.out_val + size_t(chunks) * Q * K);
  }
  dim3 grid(q_tiles, chunks);
  if (T == 1) sim_tile_kernel<0, true><<<grid, kSimThreads, 0, stream>>>(p);
  else        sim_tile_kernel<0, false><<<grid, kSimThreads, 0, stream>>>(p);
  SPN_CUDA_CHECK(cudaGetLastError());
  if (chunks > 1) return merge_topk(p.out_val, p.out_idx, chunks, Q, K, out_val, out_idx, stream);
  return SPN_OK;
}

int merge_topk(const float* vals, const int* idx, int R, int Q, int K, float* out_val, int* out_idx,
               cudaStream_t stream) {
  SPN_REQUIRE(R >= 1 && Q > 0 && K >= 1 && K <= 64, SPN_ERR_BAD_SHAPE, "merge_topk: bad shape R=%d Q=%d K=%d", R, Q, K);
  merge_topk_kernel<<<(Q + 7) / 8, 256, 0, stream>>>(vals, idx, R, Q, K, out_val, out_idx);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int sim_matrix(const float* q, const float* g, int Q, int G, int D, int T, float* out, cudaStream_t stream) {
  int rc = check_common(q, g, Q, G, D, T);
  if (rc != SPN_OK) return rc;
  int chunks, tpc;
  const int q_tiles = plan_chunks(Q, (long long)G * T, &chunks, &tpc);
  SimParams p{};
  p.q = q; p.g = g; p.Q = Q; p.G = G; p.D = D; p.T = T; p.K = 1; p.col_base = 0;
  p.tiles_per_chunk = tpc; p.out_val = out;
  dim3 grid(q_tiles, chunks);
  if (T == 1) sim_tile_kernel<2, true><<<grid, kSimThreads, 0, stream>>>(p);
  else        sim_tile_kernel<2, false><<<grid, kSimThreads, 0, stream>>>(p);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int sim_target_score(const float* q, const float* g, int Q, int G, int D, int T, const int* target, int col_base,
                     float* score, cudaStream_t stream) {
  int rc = check_common(q, g, Q, G, D, T);
  if (rc != SPN_OK) return rc;
  if (D % 64 == 0) return sim_target_score_warp(q, g, Q, G, D, T, target, col_base, score, stream);  // same chain, staged
  target_score_kernel<<<(Q + 63) / 64, 64, 0, stream>>>(q, g, Q, G, D, T, target, col_base, score);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int sim_rank(const float* q, const float* g, int Q, int G, int D, int T, const int* target, const float* tscore,
             const int* exclude, const int* group, int M, int col_base, int* rank, int* group_rank,
             cudaStream_t stream) {
  int rc = check_common(q, g, Q, G, D, T);
  if (rc != SPN_OK) return rc;
  SPN_REQUIRE(target && tscore && rank, SPN_ERR_BAD_ARG, "sim_rank needs target, target_score and rank");
  SPN_REQUIRE(!group || (M >= 1 && M <= 32 && group_rank), SPN_ERR_BAD_SHAPE, "sim_rank: group needs 1 <= M <= 32 and group_rank");
  int chunks, tpc;
  const int q_tiles = plan_chunks(Q, (long long)G * T, &chunks, &tpc);
  SimParams p{};
  p.q = q; p.g = g; p.Q = Q; p.G = G; p.D = D; p.T = T; p.K = 1; p.col_base = col_base;
  p.tiles_per_chunk = tpc; p.exclude = exclude; p.target = target; p.tscore = tscore; p.group = group; p.M = M;
  p.rank = rank; p.group_rank = group_rank;
  dim3 grid(q_tiles, chunks);
  if (T == 1) sim_tile_kernel<1, true><<<grid, kSimThreads, 0, stream>>>(p);
  else        sim_tile_kernel<1, false><<<grid, kSimThreads, 0, stream>>>(p);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
