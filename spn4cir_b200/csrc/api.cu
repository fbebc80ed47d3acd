// extern "C" surface of libspn_b200 (include/spn_b200.h): argument validation and dispatch only.
#include "bank_ce.cuh"
#include "sim_rank.cuh"

using namespace spn;

static inline cudaStream_t as_stream(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int spn_workspace_bytes(int op, int B, int64_t N, int D, int T, int K, size_t* out_bytes) {
  SPN_REQUIRE(out_bytes != nullptr, SPN_ERR_BAD_ARG, "spn_workspace_bytes: out_bytes is NULL");
  switch (op) {
    case SPN_OP_BANK_CE_FWD:
    case SPN_OP_BANK_CE_BWD: {
      BankCEPlan plan;
      int rc = bank_ce_plan(op == SPN_OP_BANK_CE_FWD ? 0 : 1, B, N, D, T, &plan);
      if (rc != SPN_OK) return rc;
      *out_bytes = plan.ws_bytes;
      return SPN_OK;
    }
    case SPN_OP_SIM_TOPK:
      SPN_REQUIRE(B > 0 && N > 0 && N < (1ll << 31) && K >= 1 && K <= 64 && (T == 1 || T == 32), SPN_ERR_BAD_SHAPE,
                  "spn_workspace_bytes(SIM_TOPK): bad shape Q=%d G=%lld T=%d K=%d", B, (long long)N, T, K);
      *out_bytes = sim_topk_ws_bytes(B, int(N), T, K);
      return SPN_OK;
    case SPN_OP_SIM_RANK:
      *out_bytes = 256;
      return SPN_OK;
    case SPN_OP_SIM_TC:
      SPN_REQUIRE(B > 0 && N > 0 && N * T < (1ll << 31) && D >= 64 && D % 64 == 0 && (T == 1 || T == 32), SPN_ERR_BAD_SHAPE,
                  "spn_workspace_bytes(SIM_TC): bad shape Q=%d G=%lld D=%d T=%d", B, (long long)N, D, T);
      *out_bytes = sim_tc_ws_bytes(B, int(N), D, T);
      return SPN_OK;
    case SPN_OP_BANK_CE_STEP:
      SPN_REQUIRE(B > 0 && D >= 64 && D <= 768 && D % 64 == 0, SPN_ERR_BAD_SHAPE, "spn_workspace_bytes(BANK_CE_STEP): bad shape B=%d D=%d", B, D);
      *out_bytes = bank_ce_step_ws_bytes(B, N, D, T);
      return SPN_OK;
    case SPN_OP_BANK_CE_FUSED:
      *out_bytes = (bank_ce_fused_ws_bytes(B, N, D, T) + 255) / 256 * 256;
      return SPN_OK;
    default:
      return fail(SPN_ERR_BAD_ARG, "spn_workspace_bytes: unknown op %d", op);
  }
}

int spn_cast_to_bf16(const void* src, int src_dtype, void* dst_bf16, int64_t n, void* stream) {
  SPN_REQUIRE(src && dst_bf16, SPN_ERR_BAD_ARG, "spn_cast_to_bf16: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return cast_to_bf16(src, src_dtype, dst_bf16, n, as_stream(stream));
}

int spn_query_assemble(const float* refer_bank, const int64_t* refer_idx, const float* text, int B, int D, int64_t R,
                       void* q_bf16, float* q_f32, float* inv_norm, void* stream) {
  SPN_REQUIRE(refer_bank && refer_idx && text && q_bf16, SPN_ERR_BAD_ARG, "spn_query_assemble: NULL pointer");
  SPN_REQUIRE(R > 0, SPN_ERR_BAD_SHAPE, "spn_query_assemble: the refer bank has no rows");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return query_assemble(refer_bank, reinterpret_cast<const long long*>(refer_idx), text, B, D, R, q_bf16, q_f32, inv_norm,
                        as_stream(stream));
}

int spn_query_assemble_bwd(const float* dq, const float* q_f32, const float* inv_norm, const float* gscale_dev, int B,
                           int D, float* dtext, void* stream) {
  SPN_REQUIRE(dq && q_f32 && inv_norm && dtext, SPN_ERR_BAD_ARG, "spn_query_assemble_bwd: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return query_assemble_bwd(dq, q_f32, inv_norm, gscale_dev, B, D, dtext, as_stream(stream));
}

int spn_bank_ce_fwd(const void* q_bf16, const void* bank_bf16, const int64_t* labels, float inv_tau, int B,
                    int64_t N, int D, int T, int64_t row_base, float* m, float* l, float* s_t, float* s_sum,
                    void* ws, size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q_bf16 && bank_bf16 && labels && m && l && s_t, SPN_ERR_BAD_ARG, "spn_bank_ce_fwd: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_forward(q_bf16, bank_bf16, reinterpret_cast<const long long*>(labels), inv_tau, B, N, D, T, row_base,
                         m, l, s_t, s_sum, ws, ws_bytes, as_stream(stream));
}

int spn_merge_stats(const float* m_parts, const float* l_parts, const float* st_parts, const float* ssum_parts,
                    int P, int B, float* m, float* l, float* s_t, float* s_sum, void* stream) {
  SPN_REQUIRE(m_parts && l_parts && st_parts && m && l && s_t, SPN_ERR_BAD_ARG, "spn_merge_stats: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return merge_stats(m_parts, l_parts, st_parts, ssum_parts, P, B, m, l, s_t, s_sum, as_stream(stream));
}

int spn_bank_ce_loss(const float* m, const float* l, const float* s_t, const float* s_sum, int B,
                     int64_t N_total, float label_smoothing, float* lse, float* loss_per_query, float* loss_mean,
                     void* stream) {
  SPN_REQUIRE(m && l && s_t && lse && loss_mean, SPN_ERR_BAD_ARG, "spn_bank_ce_loss: NULL pointer");
  SPN_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, SPN_ERR_BAD_ARG, "label_smoothing must be in [0, 1)");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_loss(m, l, s_t, s_sum, B, N_total, label_smoothing, lse, loss_per_query, loss_mean, as_stream(stream));
}

int spn_bank_ce_bwd(const void* q_bf16, const void* bank_bf16, const int64_t* labels, float inv_tau,
                    const float* lse, float gscale, const float* gscale_dev, float label_smoothing,
                    int64_t N_total, int B, int64_t N, int D, int T, int64_t row_base, float* dq_partial,
                    float* dlogit_dot, void* ws, size_t ws_bytes, void* stream) {
  (void)ws;
  (void)ws_bytes;
  SPN_REQUIRE(q_bf16 && bank_bf16 && labels && lse && dq_partial, SPN_ERR_BAD_ARG, "spn_bank_ce_bwd: NULL pointer");
  SPN_REQUIRE(aligned16(dq_partial), SPN_ERR_BAD_ALIGN, "spn_bank_ce_bwd: dq_partial must be 16-byte aligned");
  SPN_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, SPN_ERR_BAD_ARG, "label_smoothing must be in [0, 1)");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_backward(q_bf16, bank_bf16, reinterpret_cast<const long long*>(labels), inv_tau, lse, gscale,
                          gscale_dev, label_smoothing, N_total, B, N, D, T, row_base, dq_partial, dlogit_dot, as_stream(stream));
}

int spn_bank_stats(const void* bank_bf16, int64_t rows, int D, float* maxnorm, float* colsum, void* stream) {
  SPN_REQUIRE(bank_bf16 && maxnorm, SPN_ERR_BAD_ARG, "spn_bank_stats: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_stats(bank_bf16, rows, D, maxnorm, colsum, as_stream(stream));
}

int spn_bank_ce_prep(const void* q_src, int src_dtype, const void* bank_bf16, const int64_t* labels, float inv_tau,
                     const float* bank_maxnorm, int B, int64_t N, int D, int T, int64_t row_base, void* q_bf16,
                     float* ref, void* stream) {
  SPN_REQUIRE(q_src && bank_bf16 && labels && bank_maxnorm && q_bf16 && ref, SPN_ERR_BAD_ARG, "spn_bank_ce_prep: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_prep(q_src, src_dtype, bank_bf16, reinterpret_cast<const long long*>(labels), inv_tau, bank_maxnorm, B, N,
                      D, T, row_base, q_bf16, ref, as_stream(stream));
}

int spn_bank_ce_fused(const void* q_bf16, const void* bank_bf16, const int64_t* labels, const float* ref, float inv_tau,
                      int B, int64_t N, int D, int T, int64_t row_base, int init, float* o_acc, float* l_acc, float* s_t,
                      int64_t* lbl_row, float* es_acc, float* bank_t, void* ws, size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q_bf16 && bank_bf16 && labels && ref && o_acc && l_acc && s_t && lbl_row, SPN_ERR_BAD_ARG,
              "spn_bank_ce_fused: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_fused(q_bf16, bank_bf16, reinterpret_cast<const long long*>(labels), ref, inv_tau, B, N, D, T, row_base,
                       init, o_acc, l_acc, s_t, reinterpret_cast<long long*>(lbl_row), es_acc, bank_t, ws, ws_bytes,
                       as_stream(stream));
}

int spn_bank_ce_finalize(const float* o_acc, const float* l_acc, const float* s_t, const float* bank_t, const float* ref,
                         const float* es_acc, const void* q_bf16, const float* colsum, float inv_tau, float gscale,
                         float label_smoothing, int64_t N_total, int B, int D, float* lse, float* loss_per_query,
                         float* loss_mean, float* dq, float* dlogit_dot, void* stream) {
  SPN_REQUIRE(o_acc && l_acc && s_t && ref && lse && loss_mean, SPN_ERR_BAD_ARG, "spn_bank_ce_finalize: NULL pointer");
  SPN_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, SPN_ERR_BAD_ARG, "label_smoothing must be in [0, 1)");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_finalize(o_acc, l_acc, s_t, bank_t, nullptr, nullptr, /*zero_loss=*/1, ref, es_acc, q_bf16, colsum, inv_tau,
                          gscale, label_smoothing, N_total, B, D, lse, loss_per_query, loss_mean, dq, dlogit_dot,
                          as_stream(stream));
}

int spn_shard_pack(const float* l_acc, const float* s_t, const float* es_acc, const int64_t* lbl_row, int Bt, float* vec,
                   void* stream) {
  SPN_REQUIRE(l_acc && s_t && lbl_row && vec, SPN_ERR_BAD_ARG, "spn_shard_pack: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return shard_pack(l_acc, s_t, es_acc, reinterpret_cast<const long long*>(lbl_row), Bt, vec, as_stream(stream));
}

int spn_shard_fold(float* o_acc, const float* vec, const float* ref, const int64_t* lbl_row, const void* bank_bf16, int Bt,
                   int D, float label_smoothing, float* s_t_out, void* stream) {
  SPN_REQUIRE(o_acc && vec && ref && lbl_row && bank_bf16 && s_t_out, SPN_ERR_BAD_ARG, "spn_shard_fold: NULL pointer");
  SPN_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, SPN_ERR_BAD_ARG, "label_smoothing must be in [0, 1)");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return shard_fold(o_acc, vec, ref, reinterpret_cast<const long long*>(lbl_row), bank_bf16, Bt, D, label_smoothing,
                    s_t_out, as_stream(stream));
}

int spn_shard_home_prologue(const void* q_src, int src_dtype, const int64_t* labels, void* const* bank_peers,
                            int64_t rows_per_rank, int64_t N_total, int T, float inv_tau, const float* bank_maxnorm, int B, int D,
                            int rank, int world, void* const* qall_peers, void* const* ref_peers, void* const* label_peers,
                            float* s_t, float* ref, float* target_rows, float* o_window, float* l_window, float* es_window,
                            void* stream) {
  SPN_REQUIRE(q_src && labels && bank_peers && bank_maxnorm && qall_peers && ref_peers && label_peers && s_t && ref &&
                  target_rows && o_window && l_window, SPN_ERR_BAD_ARG, "spn_shard_home_prologue: NULL pointer");
  SPN_REQUIRE(aligned16(o_window) && aligned16(target_rows), SPN_ERR_BAD_ALIGN, "spn_shard_home_prologue: 16-byte alignment");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return shard_home_prologue(q_src, src_dtype, reinterpret_cast<const long long*>(labels), bank_peers, rows_per_rank, N_total,
                             T, inv_tau, bank_maxnorm, B, D, rank, world, qall_peers, ref_peers, label_peers, s_t, ref,
                             target_rows, o_window, l_window, es_window, as_stream(stream));
}

int spn_peer_barrier(void* const* flag_peers, int rank, int world, uint32_t epoch, void* stream) {
  SPN_REQUIRE(flag_peers, SPN_ERR_BAD_ARG, "spn_peer_barrier: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return peer_barrier(flag_peers, rank, world, epoch, as_stream(stream));
}

int spn_shard_pull(const float* const* o_peers, const float* const* l_peers, const float* const* es_peers, int rank, int world,
                   int B, int D, float* o_out, float* l_out, float* es_out, void* stream) {
  SPN_REQUIRE(o_peers && l_peers && o_out && l_out && (!es_out || es_peers), SPN_ERR_BAD_ARG, "spn_shard_pull: NULL pointer");
  SPN_REQUIRE(aligned16(o_out), SPN_ERR_BAD_ALIGN, "spn_shard_pull: o_out must be 16-byte aligned");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return shard_pull(o_peers, l_peers, es_peers, rank, world, B, D, o_out, l_out, es_out, as_stream(stream));
}

int spn_bank_ce_fused_peers(const void* q_bf16, const void* bank_bf16, const int64_t* labels, const float* ref, float inv_tau,
                            int B, int64_t N, int D, int T, int64_t row_base, int home_rows, int world,
                            float* const* o_windows, float* const* l_windows, float* const* es_windows, float* s_t,
                            int64_t* lbl_row, void* ws, size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q_bf16 && bank_bf16 && labels && ref && o_windows && l_windows && s_t && lbl_row, SPN_ERR_BAD_ARG,
              "spn_bank_ce_fused_peers: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_fused_peers(q_bf16, bank_bf16, reinterpret_cast<const long long*>(labels), ref, inv_tau, B, N, D, T, row_base,
                             home_rows, world, o_windows, l_windows, es_windows, s_t, reinterpret_cast<long long*>(lbl_row),
                             ws, ws_bytes, as_stream(stream));
}

int spn_bank_ce_step(const void* q_src, int src_dtype, const void* bank_bf16, const int64_t* labels, float inv_tau,
                     const float* bank_maxnorm, const float* colsum, float label_smoothing, int B, int64_t N, int D, int T,
                     float* lse, float* loss_per_query, float* loss_mean, float* dq, float* dlogit_dot, void* ws,
                     size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q_src && bank_bf16 && labels && bank_maxnorm && lse && loss_mean, SPN_ERR_BAD_ARG, "spn_bank_ce_step: NULL pointer");
  SPN_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, SPN_ERR_BAD_ARG, "label_smoothing must be in [0, 1)");
  SPN_REQUIRE(label_smoothing == 0.f || colsum, SPN_ERR_BAD_ARG, "label smoothing needs the bank column sums");
  SPN_REQUIRE(!dq || aligned16(dq), SPN_ERR_BAD_ALIGN, "spn_bank_ce_step: dq must be 16-byte aligned");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return bank_ce_step(q_src, src_dtype, bank_bf16, reinterpret_cast<const long long*>(labels), inv_tau, bank_maxnorm, colsum,
                      label_smoothing, B, N, D, T, lse, loss_per_query, loss_mean, dq, dlogit_dot, ws, ws_bytes,
                      as_stream(stream));
}

int spn_part_ce(const void* q_src, int src_dtype, const void* bank_bf16, const int64_t* candidates, float inv_tau, int B,
                int K1, int D, int64_t N, float* loss_per_query, float* loss_mean, float* dq, float* dlogit_dot,
                void* stream) {
  SPN_REQUIRE(q_src && bank_bf16 && candidates && loss_mean, SPN_ERR_BAD_ARG, "spn_part_ce: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return part_ce(q_src, src_dtype, bank_bf16, reinterpret_cast<const long long*>(candidates), inv_tau, B, K1, D, N,
                 loss_per_query, loss_mean, dq, dlogit_dot, as_stream(stream));
}

int spn_sim_topk(const float* q, const float* gallery, int Q, int G, int D, int T, int K, const int32_t* exclude,
                 int32_t col_base, float* out_val, int32_t* out_idx, void* ws, size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q && gallery && out_val && out_idx, SPN_ERR_BAD_ARG, "spn_sim_topk: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_topk(q, gallery, Q, G, D, T, K, exclude, col_base, out_val, out_idx, ws, ws_bytes, as_stream(stream));
}

int spn_sim_topk_after(const float* q, const float* gallery, int Q, int G, int D, int T, int K, const int32_t* exclude,
                       int32_t col_base, const float* after_val, const int32_t* after_idx, float* out_val,
                       int32_t* out_idx, void* ws, size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q && gallery && out_val && out_idx && after_val && after_idx, SPN_ERR_BAD_ARG, "spn_sim_topk_after: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_topk(q, gallery, Q, G, D, T, K, exclude, col_base, out_val, out_idx, ws, ws_bytes, as_stream(stream),
                  after_val, after_idx);
}

int spn_merge_topk(const float* vals, const int32_t* idx, int R, int Q, int K, float* out_val, int32_t* out_idx,
                   void* stream) {
  SPN_REQUIRE(vals && idx && out_val && out_idx, SPN_ERR_BAD_ARG, "spn_merge_topk: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return merge_topk(vals, idx, R, Q, K, out_val, out_idx, as_stream(stream));
}

int spn_sim_target_score(const float* q, const float* gallery, int Q, int G, int D, int T, const int32_t* target,
                         int32_t col_base, float* score, void* stream) {
  SPN_REQUIRE(q && gallery && target && score, SPN_ERR_BAD_ARG, "spn_sim_target_score: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_target_score(q, gallery, Q, G, D, T, target, col_base, score, as_stream(stream));
}

int spn_sim_rank(const float* q, const float* gallery, int Q, int G, int D, int T, const int32_t* target,
                 const float* target_score, const int32_t* exclude, const int32_t* group, int M, int32_t col_base,
                 int32_t* rank, int32_t* group_rank, void* ws, size_t ws_bytes, void* stream) {
  (void)ws;
  (void)ws_bytes;
  SPN_REQUIRE(q && gallery, SPN_ERR_BAD_ARG, "spn_sim_rank: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_rank(q, gallery, Q, G, D, T, target, target_score, exclude, group, M, col_base, rank, group_rank,
                  as_stream(stream));
}

int spn_sim_topk_tc(const float* q, const float* gallery, int Q, int G, int D, int T, int K, const int32_t* exclude,
                    int32_t col_base, float* out_val, int32_t* out_idx, int exact_values, int32_t* n_second, void* ws,
                    size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q && gallery && out_val && out_idx, SPN_ERR_BAD_ARG, "spn_sim_topk_tc: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_topk_tc(q, gallery, Q, G, D, T, K, exclude, col_base, out_val, out_idx, exact_values, n_second, ws, ws_bytes,
                     as_stream(stream));
}

int spn_sim_rank_tc(const float* q, const float* gallery, int Q, int G, int D, int T, const int32_t* target,
                    const float* target_score, const int32_t* exclude, const int32_t* group, int M, int32_t col_base,
                    int32_t* rank, int32_t* group_rank, void* ws, size_t ws_bytes, void* stream) {
  SPN_REQUIRE(q && gallery, SPN_ERR_BAD_ARG, "spn_sim_rank_tc: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_rank_tc(q, gallery, Q, G, D, T, target, target_score, exclude, group, M, col_base, rank, group_rank, ws,
                     ws_bytes, as_stream(stream));
}

int spn_debug_sim_tc_layout(int Q, int G, int D, int T, size_t* scores_offset, int* pitch, int* block_rows) {
  SPN_REQUIRE(scores_offset && pitch && block_rows && Q > 0 && G > 0 && D > 0 && T > 0, SPN_ERR_BAD_ARG,
              "spn_debug_sim_tc_layout: bad arguments");
  return sim_tc_debug_layout(Q, G, D, T, scores_offset, pitch, block_rows);
}

int spn_rank_window(const float* scores, int R, int G, int64_t pitch, int lo, int hi, int32_t* out_idx, int32_t* overflow,
                    void* stream) {
  SPN_REQUIRE(scores && out_idx && overflow, SPN_ERR_BAD_ARG, "spn_rank_window: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return rank_window(scores, R, G, pitch, lo, hi, out_idx, overflow, as_stream(stream));
}

int spn_sim_matrix(const float* q, const float* gallery, int Q, int G, int D, int T, float* out, void* stream) {
  SPN_REQUIRE(q && gallery && out, SPN_ERR_BAD_ARG, "spn_sim_matrix: NULL pointer");
  int rc = require_sm100();
  if (rc != SPN_OK) return rc;
  return sim_matrix(q, gallery, Q, G, D, T, out, as_stream(stream));
}

int spn_debug_read_timeline(unsigned long long* host_out, int n) {
  SPN_REQUIRE(host_out && n != 0 && n <= 1024 * 16 && n >= -1024 * 16, SPN_ERR_BAD_ARG, "spn_debug_read_timeline: bad arguments");
  return debug_read_timeline(host_out, n);
}

int spn_debug_umma_bench(const void* b_src, int64_t b_rows, int N, int iters, int mode, int tma_on,
                         unsigned long long* out_cycles, int grid, void* stream) {
  SPN_REQUIRE(out_cycles && iters > 0 && grid > 0, SPN_ERR_BAD_ARG, "spn_debug_umma_bench: bad arguments");
  return debug_umma_bench(b_src, b_rows, N, iters, mode, tma_on, out_cycles, grid, as_stream(stream));
}

int spn_debug_dsmem_bench(int iters, int ndst, unsigned long long* out_cycles, int grid, void* stream) {
  SPN_REQUIRE(out_cycles && iters > 0, SPN_ERR_BAD_ARG, "spn_debug_dsmem_bench: bad arguments");
  return debug_dsmem_bench(iters, ndst, out_cycles, grid, as_stream(stream));
}

int spn_debug_umma_gemm(const void* a_bf16, const void* b_bf16, float* c, int N, int K, int mode, void* stream) {
  SPN_REQUIRE(a_bf16 && b_bf16 && c, SPN_ERR_BAD_ARG, "spn_debug_umma_gemm: NULL pointer");
  return debug_umma_gemm(a_bf16, b_bf16, c, N, K, mode, as_stream(stream));
}

}  // extern "C"
