// Host entry points of the retrieval kernels (sim_rank.cu), used by api.cu.
#pragma once
#include "common.cuh"

namespace spn {

size_t sim_topk_ws_bytes(int Q, int G, int T, int K);
int sim_topk(const float* q, const float* g, int Q, int G, int D, int T, int K, const int* exclude, int col_base,
             float* out_val, int* out_idx, void* ws, size_t ws_bytes, cudaStream_t stream,
             const float* after_val = nullptr, const int* after_idx = nullptr);
int merge_topk(const float* vals, const int* idx, int R, int Q, int K, float* out_val, int* out_idx,
               cudaStream_t stream);
int sim_matrix(const float* q, const float* g, int Q, int G, int D, int T, float* out, cudaStream_t stream);
int sim_target_score(const float* q, const float* g, int Q, int G, int D, int T, const int* target, int col_base,
                     float* score, cudaStream_t stream);
int sim_rank(const float* q, const float* g, int Q, int G, int D, int T, const int* target, const float* tscore,
             const int* exclude, const int* group, int M, int col_base, int* rank, int* group_rank,
             cudaStream_t stream);

// rank-window selection (rank_window.cu): columns with rank in [lo, hi) of every row of a score block, in rank order
int rank_window(const float* scores, int R, int G, long long pitch, int lo, int hi, int32_t* out, int* overflow,
                cudaStream_t stream);

// tensor-core pruned variants (sim_tc.cu): same results as sim_topk / sim_rank
size_t sim_tc_ws_bytes(int Q, int G, int D, int T);
int sim_topk_tc(const float* q, const float* g, int Q, int G, int D, int T, int K, const int* exclude, int col_base,
                float* out_val, int* out_idx, int exact_values, int* n_second, void* ws, size_t ws_bytes,
                cudaStream_t stream);
int sim_rank_tc(const float* q, const float* g, int Q, int G, int D, int T, const int* target, const float* tscore,
                const int* exclude, const int* group, int M, int col_base, int* rank, int* group_rank, void* ws,
                size_t ws_bytes, cudaStream_t stream);
int sim_target_score_warp(const float* q, const float* g, int Q, int G, int D, int T, const int* target, int col_base,
                          float* score, cudaStream_t stream);
int sim_tc_debug_layout(int Q, int G, int D, int T, size_t* scores_offset, int* pitch, int* block_rows);

}  // namespace spn
