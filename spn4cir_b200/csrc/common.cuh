// Host-side helpers shared by the C-ABI translation units: thread-local error text, CUDA error
// mapping, TMA tensor-map construction (driver entry point fetched through the runtime, so the
// library does not link libcuda directly).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/spn_b200.h"
#include "../../include/spn_b200_debug.h"

namespace spn {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define SPN_CUDA_CHECK(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::spn::fail(SPN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                                \
  } while (0)

#define SPN_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return ::spn::fail((code), __VA_ARGS__); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Number of SMs / opt-in shared memory of the current device (cached per device).
int device_sm_count();
int device_max_smem_optin();
// SPN_OK if the current device is sm_100; SPN_ERR_CUDA otherwise (no fallback path exists).
int require_sm100();

// 2-D bf16 row-major tensor [rows, cols] -> tensor map with box [box_rows, 64 cols], 128B swizzle.
// Out-of-bounds rows/cols are zero-filled by the TMA unit.
int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
// bf16 [rows, cols] as a 3-D map (64-col block, row, column): box = [2 blocks][box_rows][64 cols], 128B swizzle -- one
// TMA instruction fills two consecutive [box_rows x 64] operand tiles.
int make_tmap_bf16_3d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
// bf16 [rows, cols] with an un-swizzled box [box_rows, 256 cols]: for cp.async.bulk.prefetch.tensor (L2) only.
int make_tmap_bf16_2d_prefetch(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
// 2-D fp32 row-major tensor [rows, cols] -> tensor map with box [box_rows, 32 cols], 128B swizzle (TMA reduce-add target).
int make_tmap_f32_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows);

}  // namespace spn
