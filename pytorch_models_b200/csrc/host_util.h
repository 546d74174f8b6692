// Host-side helpers shared by the C-ABI entry points: error reporting, CUtensorMap construction (cached), the
// per-device SM count and the abort word that bounded device-side waits raise (ptx.cuh: mbar_wait).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

// Thread-local last-error text, returned by b200enc_last_error().
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define B200_CHECK_ARG(cond, ...)                      \
  do {                                                 \
    if (!(cond)) return b200::set_error(-1, __VA_ARGS__); \
  } while (0)

#define B200_CUDA(call)                                                                               \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) return b200::set_error(int(e__), "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// Row-major bf16 tensor viewed as up to 3 dims (inner, rows, batches); strides in elements.
// box = (box_inner, box_rows, 1). swizzle_bytes in {0, 32, 64, 128}.
// Encoded maps are kept in a process-wide cache keyed by every argument (pointer, shape, strides, box, swizzle): a
// CUtensorMap is a pure function of those, so the cache never goes stale and a forward pass that reuses its buffers
// (the caching allocator hands the same blocks back) encodes each map once.
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                   uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows,
                   int swizzle_bytes);

// General form (rank <= 5): dims / box in elements, strides[i] = byte stride of dim i+1. Same cache.
int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes, int elem_bytes = 2);

// The same 2-D / 3-D view for 1-byte elements (e4m3 operands of the optional FP8 linears); strides in elements.
int make_tmap_u8(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches, uint64_t row_stride,
                 uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes);

struct TmapCacheStats {
  unsigned long long hits, misses;
};
TmapCacheStats tmap_cache_stats();

// SM count of the CURRENT device (cached per device ordinal).
int sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device) instead of once per launch.
int ensure_dynamic_smem(const void* kernel, int bytes);

// Programmatic dependent launch for the tensor-core kernels (GEMM, attention): on unless B200ENC_PDL=0 in the environment.
bool pdl_enabled();

// The abort word: one unsigned int in mapped, portable host memory (readable by the host without synchronising,
// writable by every device). 0 = healthy. Device-side waits that exceed their time limit store a non-zero code; every
// entry point calls check_abort() first and refuses to enqueue more work once it is set.
unsigned int* abort_word();  // device-usable pointer (UVA); nullptr if the allocation failed
int check_abort(const char* who);
unsigned int read_abort_word(bool clear);

}  // namespace b200
