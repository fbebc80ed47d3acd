/* spn_b200.h -- C ABI of the B200-native SPN4CIR hot path.
 *
 * The reference (BUAADreamer/SPN4CIR) is pure Python/PyTorch and has no FFI; the "operator API" of
 * its hot path is three Python call shapes.  This header is the C boundary a
 * torch-free host (ctypes, cffi, a C++ trainer) binds instead.  Each entry point cites the reference
 * lines it replaces.  The Python mirror of the reference call sites lives in spn4cir_b200/compat.py.
 *
 * Conventions
 *  - every function returns SPN_OK (0) or a negative spn_status; spn_last_error() gives a
 *    thread-local human readable message; nothing throws.
 *  - all data pointers are DEVICE pointers, 16-byte aligned, row-major, rows contiguous.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*); the library never
 *    allocates device memory: the caller owns outputs and the workspace (spn_workspace_bytes).
 *  - bank-CE operands are bf16 (the arithmetic type of the path: bf16 x bf16 -> fp32 on tcgen05);
 *    use spn_cast_to_bf16 as the prologue for fp32/fp16 tensors.  Retrieval operands are fp32.
 *  - "scaled logit" s[b][n] = inv_tau * <q[b], bank[n]>; with T > 1 (BLIP-2) a bank entry is T
 *    consecutive token rows and s[b][n] = inv_tau * max_t <q[b], bank[n*T + t]>.
 *  - there is no CPU fallback: on a machine without an sm_100 device every compute call fails with
 *    SPN_ERR_CUDA.
 */
#ifndef SPN_B200_H_
#define SPN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum spn_status {
  SPN_OK = 0,
  SPN_ERR_BAD_SHAPE = -1,
  SPN_ERR_BAD_ALIGN = -2,
  SPN_ERR_BAD_DTYPE = -3,
  SPN_ERR_WORKSPACE = -4,
  SPN_ERR_CUDA = -5,
  SPN_ERR_NCCL = -6,
  SPN_ERR_BAD_ARG = -7
} spn_status;

enum spn_op {
  SPN_OP_BANK_CE_FWD = 1,
  SPN_OP_BANK_CE_BWD = 2,
  SPN_OP_SIM_TOPK = 3,
  SPN_OP_SIM_RANK = 4,
  SPN_OP_BANK_CE_STEP = 5,
  SPN_OP_BANK_CE_FUSED = 6,
  SPN_OP_SIM_TC = 7 /* spn_sim_topk_tc / spn_sim_rank_tc (K is ignored) */
};

enum spn_dtype { SPN_F32 = 0, SPN_F16 = 1, SPN_BF16 = 2 };

/* ABI version, bumped on any signature change. */
int spn_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char* spn_last_error(void);
/* 1 if device `dev` is an sm_100 part this library can run on, else 0 (never fails). */
int spn_device_supported(int dev);

/* Workspace size for one call of `op` with these shapes (bytes, multiple of 256). */
int spn_workspace_bytes(int op, int B, int64_t N, int D, int T, int K, size_t* out_bytes);

/* Prologue: cast n elements of `src` (spn_dtype) to bf16.  Replaces the implicit autocast down-cast
 * of clip4cir/train_negplus.py:110 and the per-step `.to(device)` of models_negplus.py:138 (the bank
 * is cast once and stays resident). */
int spn_cast_to_bf16(const void* src, int src_dtype, void* dst_bf16, int64_t n, void* stream);

/* Prologue: q[b] = normalize(refer_bank[refer_idx[b]] + text[b]) -> bf16 (and fp32 copy
 * + inverse norm for the normalize-backward).  Replaces clip4cir/models_negplus.py:133-137.
 * refer_bank fp32 [R, D]; text fp32 [B, D]; q_f32 / inv_norm nullable.  An index outside [0, R) (an IndexError in the
 * reference) yields a NaN row -- never an out-of-bounds read. */
int spn_query_assemble(const float* refer_bank, const int64_t* refer_idx, const float* text, int B, int D, int64_t R,
                       void* q_bf16, float* q_f32, float* inv_norm, void* stream);

/* Backward of that normalisation (autograd through F.normalize, clip4cir/models_negplus.py:137; the gathered reference
 * rows are detached, :133-135, so only text receives a gradient):
 *   dtext[b] = g * inv_norm[b] * (dq[b] - q[b] <q[b], dq[b]>),  g = *gscale_dev (device scalar, nullable -> 1). */
int spn_query_assemble_bwd(const float* dq, const float* q_f32, const float* inv_norm, const float* gscale_dev, int B,
                           int D, float* dtext, void* stream);

/* Forward statistics of CrossEntropy((q @ bank^T) * inv_tau, labels) over one bank shard.
 * Replaces the logits + log-softmax of CIRPlus.infonce_loss (clip4cir/models_negplus.py:150-154,
 * zscir/models_bank.py:130-134, blip4cir/models.py:117-121, tgcir/models.py:289-293) and, with
 * T = 32, the per-sample loop of Blip2QformerCirAlignPrompt.forward_stage2
 * (blip24cir/lavis/models/blip2_models/blip2_qformer_cir_align_prompt.py:254-268).
 *   q      bf16 [B, D]            bank  bf16 [N*T, D]  (N entries of this shard)
 *   labels int64 [B], GLOBAL entry index; entries [row_base, row_base + N) live in this shard
 * Outputs (fp32 [B], natural-log units):
 *   m = max_n s, l = sum_n exp(s - m), s_t = s[label] or -inf when the label is not in this shard,
 *   s_sum (nullable) = sum_n s  (only needed for label smoothing, zscir/models_bank.py:35).
 * The B x N logit matrix is never written to memory. */
int spn_bank_ce_fwd(const void* q_bf16, const void* bank_bf16, const int64_t* labels, float inv_tau, int B,
                    int64_t N, int D, int T, int64_t row_base, float* m, float* l, float* s_t, float* s_sum,
                    void* ws, size_t ws_bytes, void* stream);

/* Merge P partial statistics laid out [P][B] (chunks of one shard, or the all-gathered shards of R
 * ranks) into one [B] set: M = max m_p, L = sum l_p exp(m_p - M), s_t = max, s_sum = sum. */
int spn_merge_stats(const float* m_parts, const float* l_parts, const float* st_parts, const float* ssum_parts,
                    int P, int B, float* m, float* l, float* s_t, float* s_sum, void* stream);

/* lse[b] = m + log(l); loss_b = lse - (1-eps) s_t - eps * s_sum / N_total; loss_mean = mean_b.
 * Replaces nn.CrossEntropyLoss(reduction='mean') (clip4cir/models_negplus.py:33,154).
 * loss_per_query nullable; s_sum may be NULL when eps == 0. */
int spn_bank_ce_loss(const float* m, const float* l, const float* s_t, const float* s_sum, int B,
                     int64_t N_total, float label_smoothing, float* lse, float* loss_per_query, float* loss_mean,
                     void* stream);

/* Backward over one bank shard, recomputing the softmax tile by tile from the GLOBAL lse:
 *   dlogit[b][n] = softmax(s)[b][n] - (1-eps) [n == label_b] - eps / N_total
 *   dq_partial[b] += g * sum_n dlogit[b][n] * bank[n]           (g = gscale * (gscale_dev ? *gscale_dev : 1);
 *                                                                 the caller passes gscale = inv_tau / B_total and
 *                                                                 the upstream scalar grad as a device pointer,
 *                                                                 so no host sync is needed)
 *   dlogit_dot[b] += sum_n dlogit[b][n] * s[b][n]               (nullable; gives d loss / d tau)
 * Replaces autograd through clip4cir/models_negplus.py:151-154 (and the T = 32 loop, where the
 * gradient flows through the arg-max token only).  dq_partial fp32 [B, D] must be zeroed by the
 * caller before the first shard/call; the bank receives no gradient (it is detached,
 * models_negplus.py:138). */
int spn_bank_ce_bwd(const void* q_bf16, const void* bank_bf16, const int64_t* labels, float inv_tau,
                    const float* lse, float gscale, const float* gscale_dev, float label_smoothing,
                    int64_t N_total, int B, int64_t N, int D, int T, int64_t row_base, float* dq_partial,
                    float* dlogit_dot, void* ws, size_t ws_bytes, void* stream);

/* ---- single-pass training path (forward statistics AND the unnormalised gradient in one sweep of the bank) ----
 * The soft-max is taken against a STATIC per-row reference ref[b] >= (approximately) max_n s[b][n]:
 *     ref = min( inv_tau |q_b| max_n |bank_n|  (Cauchy-Schwarz, can never overflow),  s_t + 30 )
 * bf16 / fp32 share an 8-bit exponent, so P~ = exp(s - ref) needs no running-max rescaling and the partial
 * sums of different CTAs / shards / ranks simply ADD (no merge pass).  Replaces, for a training step, the pair
 * spn_bank_ce_fwd + spn_bank_ce_bwd (reference: clip4cir/models_negplus.py:150-154 + autograd, one call site
 * clip4cir/train_negplus.py:110-121). */

/* One-off per bank: maxnorm[0] = max_n |bank_n| (device float), colsum (nullable, fp32 [D]) = sum_n bank_n
 * (only needed for label smoothing). */
int spn_bank_stats(const void* bank_bf16, int64_t rows, int D, float* maxnorm, float* colsum, void* stream);

/* Per step prologue: cast q (spn_dtype) to bf16 and compute ref[b] (natural units) for this shard.  A rank that
 * does not own label b uses the Cauchy-Schwarz bound alone; ranks must agree on ref (min all-reduce) before
 * spn_bank_ce_fused when the bank is sharded. */
int spn_bank_ce_prep(const void* q_src, int src_dtype, const void* bank_bf16, const int64_t* labels, float inv_tau,
                     const float* bank_maxnorm, int B, int64_t N, int D, int T, int64_t row_base, void* q_bf16,
                     float* ref, void* stream);

/* One sweep over the shard.  Outputs are ACCUMULATED (sum over shards / ranks; init != 0 zeroes them first):
 *   o_acc [B, D] += sum_{n != label} exp(s - ref) bank_n        l_acc [B] += sum_{n != label} exp(s - ref)
 *   es_acc [B] (nullable) += sum_{n != label} exp(s - ref) s    s_t [B] = s[label] (written by the owner, else -inf)
 *   lbl_row [B] = LOCAL token row of the target (arg-max token when T = 32; -1 if not in this shard)
 *   bank_t [B, D] (nullable) = fp32 copy of those rows (zeros when not in this shard) for spn_bank_ce_finalize. */
int spn_bank_ce_fused(const void* q_bf16, const void* bank_bf16, const int64_t* labels, const float* ref, float inv_tau,
                      int B, int64_t N, int D, int T, int64_t row_base, int init, float* o_acc, float* l_acc, float* s_t,
                      int64_t* lbl_row, float* es_acc, float* bank_t, void* ws, size_t ws_bytes, void* stream);

/* Combine the (globally reduced) sums:  L = l_acc + exp(s_t - ref), lse = ref + log L,
 *   loss_b = lse - (1-eps) s_t - eps sum_n s_n / N,   loss_mean = mean_b loss_b,
 *   dq[b]  = gscale * [ (o_acc - l_acc bank_t) / L + eps bank_t - (eps / N) colsum ]      (gscale = inv_tau / B)
 *   dlogit_dot[b] = sum_n dlogit[b][n] s[b][n]  (nullable; needs es_acc).
 * (1 - p_t) = l_acc / L is formed without cancellation.  dq is the gradient for an upstream gradient of 1.
 * bank_t == NULL: the target-row terms were already folded into o_acc by spn_shard_fold (row-sharded bank). */
int spn_bank_ce_finalize(const float* o_acc, const float* l_acc, const float* s_t, const float* bank_t, const float* ref,
                         const float* es_acc, const void* q_bf16, const float* colsum, float inv_tau, float gscale,
                         float label_smoothing, int64_t N_total, int B, int D, float* lse, float* loss_per_query,
                         float* loss_mean, float* dq, float* dlogit_dot, void* stream);

/* ---- row-sharded bank (reference loss over ALL rows: zscir/models_bank.py:90-101,130-134) ----
 * Every rank sweeps its N / R rows for all Bt = R * B queries of the step (spn_bank_ce_fused with row_base); the partial
 * sums are additive.  spn_shard_pack lays the per-query scalars out as vec [4, Bt] = (l_acc, s_t of the owner else 0,
 * es_acc or 0, owned in {0, 1}) so that ONE sum all-reduce merges them.  After it, spn_shard_fold lets the rank that owns
 * label b add the target-row terms of the gradient to its partial,
 *     o_acc[b] += (-l_b + eps L_b) bank[lbl_row[b]],   L_b = l_b + exp(s_t - ref_b)      (vec = the all-reduced values)
 * and writes s_t_out[b] = s_t (or -inf when no rank owns the label: the loss is +inf, as on one GPU).  One
 * reduce-scatter(sum) of o_acc then gives every rank the rows of its own queries; spn_bank_ce_finalize(bank_t = NULL)
 * finishes them.  Host plumbing (torch.distributed / NCCL): spn4cir_b200/dist.py. */
int spn_shard_pack(const float* l_acc, const float* s_t, const float* es_acc, const int64_t* lbl_row, int Bt, float* vec,
                   void* stream);
int spn_shard_fold(float* o_acc, const float* vec, const float* ref, const int64_t* lbl_row, const void* bank_bf16, int Bt,
                   int D, float label_smoothing, float* s_t_out, void* stream);

/* ---- the same exchange WITHOUT NCCL: kernels that read and write the other ranks' memory over NVLink / NVSwitch ----
 * All windows live in peer-mapped symmetric memory (allocated and exchanged by the host side, spn4cir_b200/dist.py:
 * SymmExchange over torch.distributed._symmetric_memory); every `void* const*` / `float* const*` argument is a HOST array of
 * `world` DEVICE pointers (entry r = rank r's buffer as mapped into THIS process, own rank included).  One step is
 *   spn_shard_home_prologue -> spn_peer_barrier -> spn_bank_ce_fused_peers -> spn_peer_barrier -> spn_bank_ce_finalize.
 *
 * spn_shard_home_prologue: for each of this rank's B queries -- cast to bf16 and STORE the row into every rank's query window
 * qall [world * B, D] at row rank * B + b (the all-gather); LOAD the target entry's T rows from the bank shard of the rank
 * that owns the label (owner = label / rows_per_rank) and form s_t (exact fp32 chain, max over tokens) and the final static
 * soft-max reference ref = min(bound, max(s_t + 30, bound - 64)) -- identical on every rank by construction, so no MIN
 * all-reduce; store ref / label into every rank's metadata windows; keep s_t, ref and the fp32 target row (arg-max token)
 * locally for the finalize step; zero this rank's accumulation windows o_window [B, D], l_window [B], es_window [B]
 * (nullable).  A label outside [0, N_total) gives s_t = -inf (loss +inf, as on one GPU). */
int spn_shard_home_prologue(const void* q_src, int src_dtype, const int64_t* labels, void* const* bank_peers,
                            int64_t rows_per_rank, int64_t N_total, int T, float inv_tau, const float* bank_maxnorm, int B, int D,
                            int rank, int world, void* const* qall_peers, void* const* ref_peers, void* const* label_peers,
                            float* s_t, float* ref, float* target_rows, float* o_window, float* l_window, float* es_window,
                            void* stream);
/* Barrier across the ranks on `stream`: flag_peers[r] = rank r's uint32 [world] flag array; rank i stores `epoch` into
 * slot i of every peer (release, system scope) and waits until its own slots all reached `epoch` (acquire).  `epoch` must
 * grow by one per call on every rank.  A peer that never arrives fails the launch after ~10 s instead of hanging the GPU. */
int spn_peer_barrier(void* const* flag_peers, int rank, int world, uint32_t epoch, void* stream);
/* spn_bank_ce_fused with the reduce-scatter folded into the kernel: queries [h * home_rows, (h + 1) * home_rows) of q_bf16
 * (= the query window) belong to rank h, and their partial sums are reduce-added into o_windows[h] [home_rows, D] (TMA
 * reduce-add over NVLink from the sweep's epilogue), l_windows[h] and es_windows[h] (nullable) [home_rows].  home_rows must
 * be a multiple of 128 and B = world * home_rows an even number of 128-query tiles.  s_t / lbl_row [B] are scratch. */
int spn_bank_ce_fused_peers(const void* q_bf16, const void* bank_bf16, const int64_t* labels, const float* ref, float inv_tau,
                            int B, int64_t N, int D, int T, int64_t row_base, int home_rows, int world,
                            float* const* o_windows, float* const* l_windows, float* const* es_windows, float* s_t,
                            int64_t* lbl_row, void* ws, size_t ws_bytes, void* stream);
/* The "pull" form of the same reduce-scatter (fewer NVLink bytes when many CTA pairs share a query tile: small batches):
 *   spn_shard_home_prologue -> spn_peer_barrier -> spn_bank_ce_fused(init = 1) into this rank's OWN peer-readable buffers
 *   o_all [world * B, D], l_all, es_all [world * B] -> spn_peer_barrier -> spn_shard_pull -> spn_bank_ce_finalize.
 * spn_shard_pull: o_out[b] = sum_r o_peers[r][rank * B + b] (r = 0 .. world - 1 in that order: loads over NVLink, no atomics,
 * a deterministic cross-rank sum), likewise l_out and es_out (nullable, with es_peers).  Reference anchor as above
 * (zscir/models_bank.py:90-101,130-134: one loss over ALL bank rows). */
int spn_shard_pull(const float* const* o_peers, const float* const* l_peers, const float* const* es_peers, int rank, int world,
                   int B, int D, float* o_out, float* l_out, float* es_out, void* stream);

/* One whole training step on one GPU (un-sharded bank, row_base = 0) in a single call:
 * spn_bank_ce_prep (+ zeroing of the accumulators) -> spn_bank_ce_fused -> spn_bank_ce_finalize (the target row is read
 * from the resident bank).  Three kernel launches; every intermediate lives in `ws`
 * (spn_workspace_bytes(SPN_OP_BANK_CE_STEP, B, N, D, T, 0), 256-byte aligned).  Replaces, per step, the forward
 * clip4cir/models_negplus.py:150-154 and its autograd backward (clip4cir/train_negplus.py:110-121): loss_mean is the
 * loss, dq (nullable) the gradient w.r.t. the UN-cast q for an upstream gradient of 1, dlogit_dot (nullable) gives
 * d loss / d tau as for spn_bank_ce_finalize. */
int spn_bank_ce_step(const void* q_src, int src_dtype, const void* bank_bf16, const int64_t* labels, float inv_tau,
                     const float* bank_maxnorm, const float* colsum, float label_smoothing, int B, int64_t N, int D, int T,
                     float* lse, float* loss_per_query, float* loss_mean, float* dq, float* dlogit_dot, void* ws,
                     size_t ws_bytes, void* stream);

/* Sampled-negatives loss: query b scores only its own candidate rows candidates[b][0 .. K1) of the bank
 * (column 0 = the positive), loss_b = CE(inv_tau * <q_b, bank[cand]>, label 0), loss_mean = mean_b.  Replaces
 * CIRPlus.part_infonce_loss (clip4cir/models.py:115-128; zscir/models_bank.py:103-122) for candidate ids drawn by the
 * caller (the reference draws them with random.sample, models.py:109-113).  q in `src_dtype` (rounded to bf16 inside),
 * bank bf16 [N, D], candidates int64 [B, K1] (ids outside [0, N) are masked out), dq (nullable) fp32 [B, D] =
 * d loss_mean / d q for an upstream gradient of 1, dlogit_dot (nullable) fp32 [B] as in spn_bank_ce_bwd. */
int spn_part_ce(const void* q_src, int src_dtype, const void* bank_bf16, const int64_t* candidates, float inv_tau, int B,
                int K1, int D, int64_t N, float* loss_per_query, float* loss_mean, float* dq, float* dlogit_dot,
                void* stream);

/* Fused cosine-similarity + per-row top-K over one gallery shard (descending score, ties -> lower
 * index first).  Replaces `distances = 1 - pred @ index.T; torch.argsort(distances)` followed by the
 * [:K] slices (clip4cir/validate.py:31-32,126-127; clip4cir/cirr_test_submission.py:91-92,106) and,
 * with T = 32, the max-over-tokens of Blip2QformerCirAlignPrompt.inference (...align_prompt.py:311-319).
 *   q fp32 [Q, D]; gallery fp32 [G*T, D]; exclude (nullable) int32 [Q] GLOBAL gallery index removed
 *   from row b before ranking (the reference image, validate.py:131-134); col_base = global index of
 *   this shard's first entry.  out_val fp32 [Q, K], out_idx int32 [Q, K] (global; -1 / -inf padding
 *   when fewer than K candidates exist). */
int spn_sim_topk(const float* q, const float* gallery, int Q, int G, int D, int T, int K, const int32_t* exclude,
                 int32_t col_base, float* out_val, int32_t* out_idx, void* ws, size_t ws_bytes, void* stream);

/* Next page of a deep top-K: as spn_sim_topk, but only entries that sort strictly AFTER the cursor
 * (after_val[b], after_idx[b]) -- normally the last entry of the previous page -- are eligible.  Scores are
 * bit-identical from call to call (one fp32 accumulator per score, ascending k), so consecutive pages tile the full
 * ranking without gaps or repeats.  Used for the K = 1000 hard-negative lists of the pair-mining step
 * (zscir/srm_utils.py:147-150,177-182: `argsort(sims, descending=True)[...][:K]` with the diagonal removed). */
int spn_sim_topk_after(const float* q, const float* gallery, int Q, int G, int D, int T, int K, const int32_t* exclude,
                       int32_t col_base, const float* after_val, const int32_t* after_idx, float* out_val,
                       int32_t* out_idx, void* ws, size_t ws_bytes, void* stream);

/* R-way merge of per-shard top-K lists laid out [R][Q][K] into one [Q][K] (same ordering rule). */
int spn_merge_topk(const float* vals, const int32_t* idx, int R, int Q, int K, float* out_val, int32_t* out_idx,
                   void* stream);

/* score[b] = max_t <q[b], gallery[(target[b]-col_base)*T + t]> when target[b] is in this shard, else
 * -inf.  First phase of the sharded rank computation. */
int spn_sim_target_score(const float* q, const float* gallery, int Q, int G, int D, int T,
                         const int32_t* target, int32_t col_base, float* score, void* stream);

/* Rank of the target = number of gallery entries (minus `exclude`) that sort before it, i.e. with a
 * larger score or an equal score and a smaller global index.  Gives Recall@K for every K without a
 * sort (clip4cir/validate.py:136-154).  target_score: fp32 [Q] from spn_sim_target_score (after a
 * max all-reduce when sharded).  group (nullable) int32 [Q, M] global indices of the CIRR subset
 * (validate.py:140-142): group_rank counts only those members.  rank / group_rank int32 [Q] are
 * ACCUMULATED (+=): zero them first, sum across shards afterwards. */
int spn_sim_rank(const float* q, const float* gallery, int Q, int G, int D, int T, const int32_t* target,
                 const float* target_score, const int32_t* exclude, const int32_t* group, int M,
                 int32_t col_base, int32_t* rank, int32_t* group_rank, void* ws, size_t ws_bytes, void* stream);

/* Tensor-core variants of spn_sim_topk / spn_sim_rank: SAME results (the same fp32 score bits, the same ordering rule),
 * obtained by pruning with a tcgen05 sweep.  q and gallery (fp32) are split into bf16 hi + lo parts, the sweep
 * accumulates S~ = Qh Gh^T + Qh Gl^T + Ql Gh^T in fp32 (|S~ - S| <= eps = 4e-5 |q| max|g|) into a [Q, G] fp32 block of the
 * workspace, and one warp per query turns the row into the answer: the 64 best entries by S~ are rescored with the exact
 * fp32 chain of spn_sim_topk (ascending-k fmaf) and re-sorted; a row whose 64th approximate score is within eps of its K-th
 * exact score is completed by a second scan that rescoring every entry within eps of the boundary -- so K <= 56.  For the
 * rank, entries farther than eps from target_score are counted from S~ and the others are rescored exactly.
 * Replaces the same reference lines as spn_sim_topk / spn_sim_rank (clip4cir/validate.py:31-32,126-127,136-154;
 * blip2_qformer_cir_align_prompt.py:311-319 with T = 32).  D % 64 == 0; ws from spn_workspace_bytes(SPN_OP_SIM_TC, Q, G,
 * D, T, 0), 256-byte aligned.  n_second (nullable, device int32): number of rows that left the fast path.
 * exact_values != 0: out_val holds the exact fp32 chain of every returned entry (bit-identical to spn_sim_topk; needed when
 * lists of different shards are merged by value) -- K gathered gallery rows per query, L2-bandwidth bound.
 * exact_values == 0: out_idx is STILL the exact ranking (entries S~ cannot order -- closer than 2 eps to a neighbour -- are
 * rescored), out_val holds S~ (within eps of the exact score); this is what the reference consumes (names of the sorted
 * indices, clip4cir/cirr_test_submission.py:94-109). */
int spn_sim_topk_tc(const float* q, const float* gallery, int Q, int G, int D, int T, int K, const int32_t* exclude,
                    int32_t col_base, float* out_val, int32_t* out_idx, int exact_values, int32_t* n_second, void* ws,
                    size_t ws_bytes, void* stream);
/* spn_sim_rank_tc: target_score may be NULL when the gallery is not sharded (every target row is in `gallery`): the kernel
 * then computes the target's exact score itself (same chain as spn_sim_target_score) and one launch is saved. */
int spn_sim_rank_tc(const float* q, const float* gallery, int Q, int G, int D, int T, const int32_t* target,
                    const float* target_score, const int32_t* exclude, const int32_t* group, int M, int32_t col_base,
                    int32_t* rank, int32_t* group_rank, void* ws, size_t ws_bytes, void* stream);
/* Compatibility mode only: the dense fp32 score matrix out[Q, G] (max over T tokens), for callers that
 * still want `pred_sim` itself (blip24cir/validate.py:30,116).  The fused top-K / rank entry points
 * above never materialise it. */
int spn_sim_matrix(const float* q, const float* gallery, int Q, int G, int D, int T, float* out, void* stream);


/* Rank-window selection: for every row r of a score block scores[R, pitch] (G valid columns), the columns
 * whose rank in the descending order (ties -> lower column) falls in [lo, hi), written in rank order to
 * out_idx[R, hi - lo].  Replaces `i2i_ranks = torch.argsort(sims_intra_i2i, descending=True)` + the slices
 * `i2i_ranks[i][rank:rank_max]` (zscir/get_cir_data.py:33-34,58-64; CIRR twin :106-109,131-137): two radix selects find the
 * two boundary scores, the band between them is compacted into shared memory and sorted there -- the row itself is never
 * sorted.  hi - lo <= 12288 per call (split wider windows); *overflow (device int32) is set to 1 when the ties of the
 * boundary values do not fit the band (the caller then sorts that block). */
int spn_rank_window(const float* scores, int R, int G, int64_t pitch, int lo, int hi, int32_t* out_idx, int32_t* overflow,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPN_B200_H_ */
