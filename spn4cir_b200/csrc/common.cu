// Error plumbing, device queries and TMA tensor-map construction for libspn_b200.
#include "common.cuh"

#include <mutex>
#include <string.h>

namespace spn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct DevInfo {
  bool init = false;
  int sms = 0, smem_optin = 0, cc_major = 0, cc_minor = 0;
};
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& d = g_dev[dev];
  if (!d.init) {
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return nullptr;
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    d.init = true;
  }
  return &d;
}

int device_sm_count() {
  const DevInfo* d = dev_info();
  return d ? d->sms : 0;
}
int device_max_smem_optin() {
  const DevInfo* d = dev_info();
  return d ? d->smem_optin : 0;
}
int require_sm100() {
  const DevInfo* d = dev_info();
  if (!d) {
    cudaGetLastError();
    return fail(SPN_ERR_CUDA, "no usable CUDA device: libspn_b200 has no CPU fallback");
  }
  if (d->cc_major != 10)
    return fail(SPN_ERR_CUDA, "device is sm_%d%d; libspn_b200 is built for sm_100a only", d->cc_major, d->cc_minor);
  return SPN_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (!aligned16(gptr)) return fail(SPN_ERR_BAD_ALIGN, "TMA source pointer %p is not 16-byte aligned", gptr);
  if ((cols * 2) % 16 != 0) return fail(SPN_ERR_BAD_ALIGN, "row pitch %llu B is not a multiple of 16", (unsigned long long)(cols * 2));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64u, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%llu cols=%llu box_rows=%u", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return SPN_OK;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (!aligned16(gptr)) return fail(SPN_ERR_BAD_ALIGN, "TMA source pointer %p is not 16-byte aligned", gptr);
  if (cols % 64 != 0) return fail(SPN_ERR_BAD_SHAPE, "3-D tensor map needs a multiple of 64 columns (got %llu)", (unsigned long long)cols);
  // [rows, cols] viewed as (64-column block k, row r, column c): one box = two consecutive 64-column blocks of box_rows rows,
  // landing in shared memory as two back-to-back [box_rows x 64] 128B-swizzled tiles (what two 2-D loads would produce)
  cuuint64_t dims[3] = {64u, rows, cols / 64};
  cuuint64_t strides[2] = {cols * 2, 128u};
  cuuint32_t box[3] = {64u, box_rows, 2u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (CUresult %d) rows=%llu cols=%llu box_rows=%u", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return SPN_OK;
}

int make_tmap_bf16_2d_prefetch(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {256u, box_rows};  // un-swizzled 512-byte rows: L2 prefetch only, never a shared-memory image
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled(prefetch) failed (CUresult %d) rows=%llu cols=%llu", (int)r,
                (unsigned long long)rows, (unsigned long long)cols);
  return SPN_OK;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (!aligned16(gptr)) return fail(SPN_ERR_BAD_ALIGN, "TMA pointer %p is not 16-byte aligned", gptr);
  if ((cols * 4) % 16 != 0) return fail(SPN_ERR_BAD_ALIGN, "row pitch %llu B is not a multiple of 16", (unsigned long long)(cols * 4));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {32u, box_rows};  // 32 fp32 = one 128-byte swizzle row
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SPN_ERR_CUDA, "cuTensorMapEncodeTiled(f32) failed (CUresult %d) rows=%llu cols=%llu box_rows=%u", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return SPN_OK;
}

}  // namespace spn

extern "C" {

int spn_version(void) { return 2; }

const char* spn_last_error(void) { return spn::g_err; }

int spn_device_supported(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

}  // extern "C"
