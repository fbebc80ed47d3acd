/* spn_b200_debug.h -- bring-up / test-only entry points of libspn_b200 (NOT part of the product ABI in spn_b200.h).
 *
 * Used by tests/ and tools/ only: descriptor self-tests, issue-rate probes, per-role cycle counters, and the workspace
 * layout query the tensor-core retrieval tests use to read S~ back.  Same conventions as spn_b200.h. */
#ifndef SPN_B200_DEBUG_H_
#define SPN_B200_DEBUG_H_

#include "spn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug / tests: where the last S~ block of a tensor-core call lives in its workspace (byte offset, row pitch in floats,
 * queries per block) so that the eps bound can be checked against the exact scores. */
int spn_debug_sim_tc_layout(int Q, int G, int D, int T, size_t* scores_offset, int* pitch, int* block_rows);

/* Debug / bring-up: one-CTA tcgen05 GEMM  C[128, N] = A[128, K] * op(B)  used by tests to pin the
 * UMMA descriptor encodings.  mode bit0: A from TMEM (else smem); bit1: B is MN-major [K, N]
 * (else K-major [N, K]).  a, b bf16; c fp32 [128, N]; N in {64,128}, K multiple of 64 (<= 256),
 * MN-major requires K == 128. */
int spn_debug_umma_gemm(const void* a_bf16, const void* b_bf16, float* c, int N, int K, int mode, void* stream);

/* Debug: raw tcgen05.mma issue-rate probe (M128 x N x K64 per iteration, SS or TS (mode bit0), optionally under
 * concurrent TMA traffic).  b_src: bf16 [b_rows, 512] device buffer with b_rows >= grid*512.  out_cycles: device
 * uint64[grid]. */
int spn_debug_umma_bench(const void* b_src, int64_t b_rows, int N, int iters, int mode, int tma_on,
                         unsigned long long* out_cycles, int grid, void* stream);

/* Debug: distributed-shared-memory store bandwidth probe (clusters of 4 CTAs; ranks 0/1 write iters x 32 KiB to
 * rank 2, and to rank 3 as well when ndst == 2).  out_cycles: device uint64[grid] (0 for receiver CTAs). */
int spn_debug_dsmem_bench(int iters, int ndst, unsigned long long* out_cycles, int grid, void* stream);

/* Debug: per-CTA role cycle counters of the last bank-CE launch (16 x uint64 per CTA); only in builds made
 * with -DSPN_TIMELINE, otherwise SPN_ERR_BAD_ARG. host_out is a HOST pointer. */
int spn_debug_read_timeline(unsigned long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* SPN_B200_DEBUG_H_ */
