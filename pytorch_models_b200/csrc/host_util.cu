#include "host_util.h"

#include <string.h>

namespace b200 {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                   uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows,
                   int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(-2, "cuTensorMapEncodeTiled driver entry point not available");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return set_error(-1, "tensor base %p not 16-byte aligned", base);
  if ((row_stride * 2) % 16 != 0) return set_error(-1, "row stride %llu elements is not a multiple of 8", (unsigned long long)row_stride);
  const uint32_t rank = batches == 0 ? 2 : 3;
  if (rank == 3 && (batch_stride * 2) % 16 != 0)
    return set_error(-1, "batch stride %llu elements is not a multiple of 8", (unsigned long long)batch_stride);
  cuuint64_t dims[3] = {inner, rows, batches == 0 ? 1 : batches};
  cuuint64_t strides[2] = {row_stride * 2, batch_stride * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(int(r),
                     "cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu batches=%llu row_stride=%llu "
                     "batch_stride=%llu box=%ux%u",
                     int(r), (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batches,
                     (unsigned long long)row_stride, (unsigned long long)batch_stride, box_inner, box_rows);
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

}  // namespace b200
