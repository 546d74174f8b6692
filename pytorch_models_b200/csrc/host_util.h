// Host-side helpers shared by the C-ABI entry points: error reporting and CUtensorMap construction.
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
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                   uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows,
                   int swizzle_bytes);

int sm_count();

}  // namespace b200
