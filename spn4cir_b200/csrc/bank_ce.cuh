// Shared declarations of the bank-CE translation units (kernel parameters, launch plan, host entry
// points used by api.cu).
#pragma once
#include "common.cuh"

namespace spn {

constexpr int kBM = 128;      // queries per CTA  (UMMA M)
constexpr int kBN = 128;      // bank rows per S tile (UMMA N of GEMM1, K extent of GEMM2)
constexpr int kBK = 64;       // bf16 elements per 128-byte swizzle row
constexpr int kUmmaK = 16;    // K per tcgen05.mma for 16-bit inputs
constexpr uint32_t kBlockBytes = kBM * kBK * 2;  // one [128 x 64] bf16 block = 16 KiB
constexpr int kMaxPeers = 8;   // ranks of one NVSwitch box
constexpr int kBankCEThreads = 320;  // TMA producer warp + MMA warp + 8 soft-max / epilogue warps

struct BankCEParams {
  int B, D;
  long long N;       // bank entries in this shard
  long long n_rows;  // N * T token rows
  int kblocks;       // D / 64
  int stages;        // ring depth
  int bps;           // k-blocks (16 KiB) per ring stage
  int n_tiles;       // ceil(n_rows / 128)
  int q_tiles;       // ceil(B / 128)
  int num_chunks;    // bank-row chunks (grid = q_tiles * num_dsplit * num_chunks)
  int num_dsplit;    // backward: splits of the output dim
  int ow_blocks;     // backward: 64-column blocks of dq per split
  int hg;            // CTA-pair schedule knob (see the MMA issuer)
  int qt_blocks;     // two-pass: leading k-blocks of the Q tile that live in TMEM (TS operand) instead of shared memory
  int pdl;           // launched with programmatic stream serialization (spn_bank_ce_step chains prep -> sweep -> finalize)
  int two_pass;      // fused CTA-pair kernel: second pass over the chunk for dq[:, X:D) (no S recompute)
  int tma_single;    // fused CTA-pair kernel, single pass (D <= 320 or D-split): TMA reduce-add epilogue instead of per-thread reds
  int stages2;       // second-pass ring depth (stages of 2 blocks)
  unsigned short* p_scratch;  // two-pass: bf16 global scratch for P~, tile-major [q_tile][bank tile][128][128]
  long long p_ld;    // its row stride in elements (n_tiles * 128)
  float scale_log2;  // inv_tau * log2(e)
  long long row_base;
  const long long* labels;
  const void* q_gmem;  // forward: bf16 [B, D] queries, staged into TMEM by the soft-max warps
  // forward partials, [num_chunks][B], log2 units
  float* part_m;
  float* part_l;
  float* part_st;
  float* part_ssum;
  // backward
  const float* lse;  // [B] natural units, global over all shards
  float gscale, eps_over_n, one_minus_eps;
  const float* gscale_dev;  // nullable: upstream scalar gradient on the device
  const float* ref;   // fused: [B] static soft-max reference, natural units
  float* l_acc;       // fused: [B] += sum_{n != label} exp(s - ref)
  float* st_out;      // fused: [B] target logit (owner CTA writes)
  long long* lbl_row_out;  // fused: [B] LOCAL token row of the target (arg-max token for T = 32)
  float* dq;          // [B, D] fp32, +=  (fused: unnormalised O, label column excluded)
  float* dlogit_dot;  // [B] fp32, += (nullable)
  // fused sweep over a ROW-SHARDED bank with the reduce-scatter folded into the epilogue (shard.cu): query tile t belongs
  // to home rank t * 128 / home_rows, and its partial sums are reduce-added straight into THAT rank's window over NVLink
  // (peer-mapped symmetric memory).  home_rows == 0: single destination (dq / l_acc / dlogit_dot above).
  int home_rows;                 // queries per home rank (multiple of 128)
  float* dq_peers[kMaxPeers];    // [home_rows, D] fp32 window of every rank (own window included)
  float* l_peers[kMaxPeers];     // [home_rows] fp32
  float* es_peers[kMaxPeers];    // [home_rows] fp32 (nullable as a group)
};

// TMA reduce-add destinations of the peer windows (two-pass kernel), passed by value as a kernel parameter
struct PeerMaps {
  CUtensorMap m[kMaxPeers];
};

struct BankCEPlan {
  int kblocks, q_tiles, n_tiles, num_chunks, num_dsplit, ow_blocks, stages, bps, grid;
  int hg;
  int cg;  // 2: CTA pairs (cta_group::2), 1: single-CTA kernel
  int two_pass, stages2, qt_blocks;  // mode 2 only
  size_t smem_bytes, ws_bytes;
};

// mode 0 = forward statistics, 1 = backward / fused with the output dim split across CTAs (S recomputed per split),
// 2 = fused two-pass (CTA pairs, D > 320: needs plan.ws_bytes of scratch for P~); falls back to the mode-1 plan when the
// shape does not pair up
int bank_ce_plan(int mode, int B, long long N, int D, int T, BankCEPlan* plan);

int bank_ce_forward(const void* q, const void* bank, const long long* labels, float inv_tau, int B, long long N,
                    int D, int T, long long row_base, float* m, float* l, float* s_t, float* s_sum, void* ws,
                    size_t ws_bytes, cudaStream_t stream);
int bank_ce_backward(const void* q, const void* bank, const long long* labels, float inv_tau, const float* lse,
                     float gscale, const float* gscale_dev, float label_smoothing, long long N_total, int B,
                     long long N, int D, int T, long long row_base, float* dq, float* dlogit_dot, cudaStream_t stream);
int bank_stats(const void* bank, long long rows, int D, float* maxnorm, float* colsum, cudaStream_t stream);
int bank_ce_prep(const void* q_src, int dtype, const void* bank, const long long* labels, float inv_tau,
                 const float* maxnorm, int B, long long N, int D, int T, long long row_base, void* q16, float* ref,
                 cudaStream_t stream);
int bank_ce_fused(const void* q, const void* bank, const long long* labels, const float* ref, float inv_tau, int B,
                  long long N, int D, int T, long long row_base, int init, float* o_acc, float* l_acc, float* s_t,
                  long long* lbl_row, float* es_acc, float* bank_t, void* ws, size_t ws_bytes, cudaStream_t stream,
                  int pdl = 0);
// the same sweep with per-home-rank destinations (peer windows); l_peers / es_peers / o_peers are HOST arrays of `world`
// device pointers
int bank_ce_fused_peers(const void* q, const void* bank, const long long* labels, const float* ref, float inv_tau, int B,
                        long long N, int D, int T, long long row_base, int home_rows, int world, float* const* o_peers,
                        float* const* l_peers, float* const* es_peers, float* s_t, long long* lbl_row, void* ws,
                        size_t ws_bytes, cudaStream_t stream);
size_t bank_ce_fused_ws_bytes(int B, long long N, int D, int T);
int bank_ce_finalize(const float* o_acc, const float* l_acc, const float* s_t, const float* bank_t, const void* bank,
                     const long long* lbl_row, int zero_loss, const float* ref, const float* es_acc, const void* q16, const float* colsum, float inv_tau, float gscale, float eps,
                     long long N_total, int B, int D, float* lse, float* loss_q, float* loss_mean, float* dq,
                     float* dlogit_dot, cudaStream_t stream, int pdl = 0);
size_t bank_ce_step_ws_bytes(int B, long long N, int D, int T);
int bank_ce_step(const void* q_src, int dtype, const void* bank, const long long* labels, float inv_tau,
                 const float* maxnorm, const float* colsum, float eps, int B, long long N, int D, int T, float* lse,
                 float* loss_q, float* loss_mean, float* dq, float* dlogit_dot, void* ws, size_t ws_bytes,
                 cudaStream_t stream);
int merge_stats(const float* pm, const float* pl, const float* pst, const float* pss, int P, int B, float* m,
                float* l, float* st, float* ss, cudaStream_t stream);
int bank_ce_loss(const float* m, const float* l, const float* st, const float* ss, int B, long long N_total,
                 float eps, float* lse, float* loss_q, float* loss_mean, cudaStream_t stream);

// row-sharded bank: pack / fold kernels around the exchange (shard.cu)
int shard_pack(const float* l_acc, const float* s_t, const float* es_acc, const long long* lbl_row, int Bt, float* vec,
               cudaStream_t stream);
int shard_fold(float* o_acc, const float* vec, const float* ref, const long long* lbl_row, const void* bank, int Bt,
               int D, float eps, float* s_t_out, cudaStream_t stream);

// NVLink exchange kernels over peer-mapped symmetric memory (shard.cu); the pointer arguments `void* const*` are HOST
// arrays of `world` device pointers (one per rank, own rank included)
int shard_home_prologue(const void* q_src, int dtype, const long long* labels, void* const* bank_peers,
                        long long rows_per_rank, long long n_total, int T, float inv_tau, const float* maxnorm, int B, int D,
                        int rank, int world, void* const* qall_peers, void* const* ref_peers, void* const* lab_peers,
                        float* s_t, float* ref_out, float* tgt_rows, float* o_win, float* l_win, float* es_win,
                        cudaStream_t stream);
int peer_barrier(void* const* flag_peers, int rank, int world, unsigned int epoch, cudaStream_t stream);
int shard_pull(const float* const* o_peers, const float* const* l_peers, const float* const* es_peers, int rank, int world, int B,
               int D, float* o_out, float* l_out, float* es_out, cudaStream_t stream);

// sampled-negatives loss (part_ce.cu)
int part_ce(const void* q_src, int dtype, const void* bank, const long long* cand, float inv_tau, int B, int K1, int D,
            long long N, float* loss_q, float* loss_mean, float* dq, float* dlogit_dot, cudaStream_t stream);

int debug_read_timeline(unsigned long long* host_out, int n);

// bring-up GEMM (umma_selftest.cu)
int debug_umma_gemm(const void* a, const void* b, float* c, int N, int K, int mode, cudaStream_t stream);

int debug_umma_bench(const void* b_src, long long b_rows, int N, int iters, int mode, int tma_on,
                     unsigned long long* out_cycles, int grid, cudaStream_t stream);

int debug_dsmem_bench(int iters, int ndst, unsigned long long* out_cycles, int grid, cudaStream_t stream);

// prologue kernels (prologue.cu)
int cast_to_bf16(const void* src, int src_dtype, void* dst, long long n, cudaStream_t stream);
int query_assemble(const float* refer_bank, const long long* refer_idx, const float* text, int B, int D, long long R,
                   void* q_bf16, float* q_f32, float* inv_norm, cudaStream_t stream);
int query_assemble_bwd(const float* dq, const float* q_f32, const float* inv_norm, const float* gscale_dev, int B, int D,
                       float* dtext, cudaStream_t stream);

}  // namespace spn
