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
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>

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

// ------------------------------------------------------------------------------------------------------------------
// NVLink exchange without NCCL (peer-mapped symmetric memory; host plumbing in spn4cir_b200/dist.py: SymmExchange).
//
//   home prologue : one warp per LOCAL query -- round q to bf16 and store the row into EVERY rank's query window
//                   (peer stores over NVLink: this is the all-gather), fetch the target row(s) straight from the bank
//                   shard of the rank that owns the label (peer loads), so the home rank itself knows s_t, the final
//                   static soft-max reference (no MIN all-reduce) and the target row for the finalize step (no second
//                   reduce-scatter); ref + label go to every rank's metadata window; the rank's own accumulation windows
//                   are zeroed for the peers' reduce-adds.
//   peer barrier  : one flag per (rank, peer) in symmetric memory, release / acquire at system scope.
//   sweep         : bank_ce_kernel with per-home destinations (bank_ce_tc.cu: home_rows) -- the reduce-scatter is the
//                   kernel's own TMA reduce-add epilogue into the home rank's window.
namespace spn {
namespace {

struct PeerPtrs {
  void* p[kMaxPeers];
};

template <typename SrcT>
__device__ __forceinline__ float ld_as_f32(const SrcT* p, size_t i);
template <>
__device__ __forceinline__ float ld_as_f32<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ld_as_f32<__half>(const __half* p, size_t i) { return __half2float(p[i]); }
template <>
__device__ __forceinline__ float ld_as_f32<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }

template <typename SrcT>
__global__ void shard_home_prologue_kernel(const SrcT* __restrict__ q_src, const long long* __restrict__ labels,
                                           const PeerPtrs bank_peers, long long rows_per_rank, long long n_total, int T,
                                           float inv_tau, const float* __restrict__ maxnorm, int B, int D, int rank,
                                           int world, const PeerPtrs qall_peers, const PeerPtrs ref_peers,
                                           const PeerPtrs lab_peers, float* __restrict__ s_t, float* __restrict__ ref_out,
                                           float* __restrict__ tgt_rows, float* __restrict__ o_win,
                                           float* __restrict__ l_win, float* __restrict__ es_win) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  // this lane's slice of the row, rounded to bf16 once: elements d = 2 lane + 64 k (+1), k < D / 64 <= 12
  float2 qv[12];
  uint32_t qbits[12];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const int d = k * 64 + lane * 2;
    qv[k] = make_float2(0.f, 0.f);
    qbits[k] = 0u;
    if (d < D) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(ld_as_f32<SrcT>(q_src, size_t(b) * D + d),
                                                     ld_as_f32<SrcT>(q_src, size_t(b) * D + d + 1));
      qbits[k] = *reinterpret_cast<const uint32_t*>(&h);
      qv[k] = __bfloat1622float2(h);
      ss = fmaf(qv[k].x, qv[k].x, fmaf(qv[k].y, qv[k].y, ss));
    }
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  // the all-gather: this row into every rank's [world * B, D] bf16 window at row rank * B + b
  const size_t grow = size_t(rank) * B + b;
  for (int r = 0; r < world; ++r) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(static_cast<__nv_bfloat16*>(qall_peers.p[r]) + grow * D);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int d = k * 64 + lane * 2;
      if (d < D) dst[d >> 1] = qbits[k];
    }
  }
  // target logit from the owner's shard (peer loads over NVLink: one round trip per token row, the values stay in
  // registers for the fp32 target row), exact fp32 chain over the bf16 values; T > 1: first arg-max token
  float r = inv_tau * sqrtf(ss) * (*maxnorm) * 1.01f + 1e-3f;
  const long long lab = labels[b];
  float best = -INFINITY;
  bool have = false;
  float2 bestv[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) bestv[k] = make_float2(0.f, 0.f);
  if (lab >= 0 && lab < n_total) {
    const int owner = int(lab / rows_per_rank);
    const long long local = lab - (long long)owner * rows_per_rank;
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(bank_peers.p[owner]) + size_t(local) * T * D;
    for (int tok = 0; tok < T; ++tok) {
      const __nv_bfloat16* row = base + size_t(tok) * D;
      float2 v[12];
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int d = k * 64 + lane * 2;
        v[k] = make_float2(0.f, 0.f);
        if (d < D) v[k] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + d));
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) acc = fmaf(qv[k].x, v[k].x, fmaf(qv[k].y, v[k].y, acc));
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (acc > best || !have) {
        best = acc;
        have = true;
#pragma unroll
        for (int k = 0; k < 12; ++k) bestv[k] = v[k];
      }
    }
    r = fminf(r, fmaxf(best * inv_tau + 30.f, r - 64.f));
  }
  // target row (fp32) for the finalize step; zeros when the label is out of range (loss = +inf, as on one GPU)
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const int d = k * 64 + lane * 2;
    if (d < D) *reinterpret_cast<float2*>(tgt_rows + size_t(b) * D + d) = bestv[k];
  }
  if (lane == 0) {
    s_t[b] = have ? best * inv_tau : -INFINITY;
    ref_out[b] = r;
    for (int rr = 0; rr < world; ++rr) {
      static_cast<float*>(ref_peers.p[rr])[grow] = r;
      static_cast<long long*>(lab_peers.p[rr])[grow] = lab;
    }
    l_win[b] = 0.f;
    if (es_win) es_win[b] = 0.f;
  }
  float4* o = reinterpret_cast<float4*>(o_win + size_t(b) * D);
  for (int d = lane; d < D / 4; d += 32) o[d] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// one warp: lane p signals peer p (its slot for this rank) and waits for peer p's signal in this rank's own flags
__global__ void peer_barrier_kernel(const PeerPtrs flag_peers, int rank, int world, unsigned int epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();  // everything this GPU wrote before the barrier (previous kernels incl. peer stores) is ordered first
  unsigned int* remote = static_cast<unsigned int*>(flag_peers.p[p]) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const unsigned int* mine = static_cast<const unsigned int*>(flag_peers.p[rank]) + p;
  const unsigned long long t0 = clock64();
  unsigned int v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (int(v - epoch) >= 0) break;
    if (clock64() - t0 > 20000000000ull) {  // ~10 s: a peer never arrived -- fail the launch instead of hanging the GPU
      printf("spn peer barrier: rank %d timed out waiting for rank %d (epoch %u, saw %u)\n", rank, p, epoch, v);
      __trap();
    }
  } while (true);
  __threadfence_system();
}


// The reduce-scatter of the "pull" exchange: every rank swept its shard for ALL world * B queries into its OWN (peer-readable)
// buffers o_all [world * B, D], l_all, es_all; the home rank of rows [rank * B, rank * B + B) now LOADS those rows from every
// rank over NVLink and sums them in the fixed order r = 0 .. world - 1 (no atomics: the cross-rank sum is deterministic).
// All loads of one element are issued before the first add (<= 8 independent 16-byte loads in flight per thread).
__global__ void __launch_bounds__(256)
shard_pull_kernel(const PeerPtrs o_peers, const PeerPtrs l_peers, const PeerPtrs es_peers, int world, int rank, int B, int D,
                  float* __restrict__ o_out, float* __restrict__ l_out, float* __restrict__ es_out) {
  const size_t n4 = size_t(B) * D / 4, base4 = size_t(rank) * B * D / 4;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    float4 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) v[r] = __ldcg(static_cast<const float4*>(o_peers.p[r]) + base4 + i);
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < kMaxPeers; ++r)
      if (r < world) {
        acc.x += v[r].x;
        acc.y += v[r].y;
        acc.z += v[r].z;
        acc.w += v[r].w;
      }
    reinterpret_cast<float4*>(o_out)[i] = acc;
  }
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      float l = 0.f, es = 0.f;
      for (int r = 0; r < world; ++r) {
        l += __ldcg(static_cast<const float*>(l_peers.p[r]) + size_t(rank) * B + b);
        if (es_out) es += __ldcg(static_cast<const float*>(es_peers.p[r]) + size_t(rank) * B + b);
      }
      l_out[b] = l;
      if (es_out) es_out[b] = es;
    }
  }
}

}  // namespace

int shard_home_prologue(const void* q_src, int dtype, const long long* labels, void* const* bank_peers,
                        long long rows_per_rank, long long n_total, int T, float inv_tau, const float* maxnorm, int B, int D,
                        int rank, int world, void* const* qall_peers, void* const* ref_peers, void* const* lab_peers,
                        float* s_t, float* ref_out, float* tgt_rows, float* o_win, float* l_win, float* es_win,
                        cudaStream_t stream) {
  SPN_REQUIRE(B > 0 && D >= 64 && D <= 768 && D % 64 == 0 && (T == 1 || T == 32), SPN_ERR_BAD_SHAPE,
              "shard_home_prologue: bad shape B=%d D=%d T=%d", B, D, T);
  SPN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && rows_per_rank > 0, SPN_ERR_BAD_ARG,
              "shard_home_prologue: bad rank / world (%d / %d)", rank, world);
  PeerPtrs bp{}, qp{}, rp{}, lp{};
  for (int r = 0; r < world; ++r) {
    SPN_REQUIRE(bank_peers[r] && qall_peers[r] && ref_peers[r] && lab_peers[r], SPN_ERR_BAD_ARG, "shard_home_prologue: NULL window of rank %d", r);
    bp.p[r] = bank_peers[r]; qp.p[r] = qall_peers[r]; rp.p[r] = ref_peers[r]; lp.p[r] = lab_peers[r];
  }
  const int grid = (B + 3) / 4;
#define SPN_LAUNCH_HOME(SrcT)                                                                                          \
  shard_home_prologue_kernel<SrcT><<<grid, 128, 0, stream>>>(static_cast<const SrcT*>(q_src), labels, bp, rows_per_rank, \
                                                             n_total, T, inv_tau, maxnorm, B, D, rank, world, qp, rp, lp, \
                                                             s_t, ref_out, tgt_rows, o_win, l_win, es_win)
  switch (dtype) {
    case SPN_F32: SPN_LAUNCH_HOME(float); break;
    case SPN_F16: SPN_LAUNCH_HOME(__half); break;
    case SPN_BF16: SPN_LAUNCH_HOME(__nv_bfloat16); break;
    default: return fail(SPN_ERR_BAD_DTYPE, "shard_home_prologue: unknown dtype %d", dtype);
  }
#undef SPN_LAUNCH_HOME
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int peer_barrier(void* const* flag_peers, int rank, int world, unsigned int epoch, cudaStream_t stream) {
  SPN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, SPN_ERR_BAD_ARG, "peer_barrier: bad rank / world");
  PeerPtrs fp{};
  for (int r = 0; r < world; ++r) {
    SPN_REQUIRE(flag_peers[r], SPN_ERR_BAD_ARG, "peer_barrier: NULL flags of rank %d", r);
    fp.p[r] = flag_peers[r];
  }
  peer_barrier_kernel<<<1, 32, 0, stream>>>(fp, rank, world, epoch);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

int shard_pull(const float* const* o_peers, const float* const* l_peers, const float* const* es_peers, int rank, int world, int B,
               int D, float* o_out, float* l_out, float* es_out, cudaStream_t stream) {
  SPN_REQUIRE(B > 0 && D > 0 && D % 4 == 0, SPN_ERR_BAD_SHAPE, "shard_pull: bad shape B=%d D=%d", B, D);
  SPN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, SPN_ERR_BAD_ARG, "shard_pull: bad rank / world");
  PeerPtrs op{}, lp{}, ep{};
  for (int r = 0; r < world; ++r) {
    SPN_REQUIRE(o_peers[r] && l_peers[r] && (!es_out || (es_peers && es_peers[r])), SPN_ERR_BAD_ARG, "shard_pull: NULL buffer of rank %d", r);
    op.p[r] = const_cast<float*>(o_peers[r]);
    lp.p[r] = const_cast<float*>(l_peers[r]);
    ep.p[r] = es_out ? const_cast<float*>(es_peers[r]) : nullptr;
  }
  const long long n4 = (long long)B * D / 4;
  const long long blocks = (n4 + 255) / 256;
  const int grid = int(blocks < 148 * 8 ? blocks : 148 * 8);
  shard_pull_kernel<<<grid, 256, 0, stream>>>(op, lp, ep, world, rank, B, D, o_out, l_out, es_out);
  SPN_CUDA_CHECK(cudaGetLastError());
  return SPN_OK;
}

}  // namespace spn
