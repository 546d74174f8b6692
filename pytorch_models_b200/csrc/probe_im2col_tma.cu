// Probe (not shipped): does cuTensorMapEncodeTiled accept the 5-D "patch" view of an NCHW bf16 image whose strides are
// NOT monotonic (dims j, i, pw, ph, n*c with byte strides 2, 2W, 2p, 2pW, 2HW), and does one cp.async.bulk.tensor.5d
// box land in shared memory as [patch][i][j] rows of 128 bytes under the 128B swizzle? Prints the verdict.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "ptx.cuh"
using namespace b200;

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int i0, int ph0, int nc, int rows, int bytes, int swz) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem), b = sb + 16384;
  if (sb & 1023u) {
    if (threadIdx.x == 0) out[0] = 0xdead;
    return;
  }
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(b, bytes);
    tma_load_5d(&tm, b, sb, 0, i0, 0, ph0, nc);
  }
  mbar_wait_plain(b, 0);
  for (int idx = threadIdx.x; idx < rows * 64; idx += blockDim.x) {
    const int r = idx / 64, e = idx % 64;
    const int chunk = swz ? (e / 8) ^ (r & 7) : e / 8;  // undo the 128B swizzle
    out[idx] = reinterpret_cast<uint16_t*>(smem + r * 128 + chunk * 16)[e % 8];
  }
}


// Variant 3: the layout the tensor core can consume WITHOUT swizzle. K-major no-swizzle operands are made of 8-row x
// 16-byte core matrices (128 contiguous bytes); a 5-D view (j8, R = plane*Hp + ph, jh, i, pw) with byte strides
// (2, 16*W*2, 16, W*2, 32) and box (8, 8, 2, 4, 16) writes exactly that: [pw][i][jh][ph%8][8 pixels].
__global__ void probe_cm(const __grid_constant__ CUtensorMap tm, uint16_t* out, int R0, int i0, int pw0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem), b = sb + 16384;
  if (threadIdx.x == 0) {
    mbar_init(b, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(b, 16384);
    tma_load_5d(&tm, b, sb, 0, R0, 0, i0, pw0);
  }
  mbar_wait_plain(b, 0);
  for (int idx = threadIdx.x; idx < 8192; idx += blockDim.x) out[idx] = reinterpret_cast<uint16_t*>(smem)[idx];
}

static int run_core_matrix_probe() {
  const int N = 2, H = 224, W = 224, p = 16, Wp = W / p, Hp = H / p;
  std::vector<uint16_t> img(size_t(N) * 3 * H * W);
  for (size_t i = 0; i < img.size(); ++i) img[i] = uint16_t(i * 2654435761u >> 16);
  uint16_t *dimg, *dout;
  cudaMalloc(&dimg, img.size() * 2);
  cudaMalloc(&dout, 8192 * 2);
  cudaMemcpy(dimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn fn = (Fn)fnp;
  CUtensorMap tm;
  cuuint64_t dims[5] = {8, cuuint64_t(N) * 3 * Hp, 2, cuuint64_t(p), cuuint64_t(Wp)};
  cuuint64_t st[4] = {cuuint64_t(16) * W * 2, 16, cuuint64_t(W) * 2, 32};
  cuuint32_t box[5] = {8, 8, 2, 4, 16};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dimg, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (j8, R, jh, i, pw): CUresult %d\n", int(r));
  if (r != CUDA_SUCCESS) return 1;
  const int plane = 4, ph0 = 8, i0 = 4, pw0 = 0;  // image 1 channel 1, patch rows 8..15 (14, 15 = next plane), pixel rows 4..7
  cudaFuncSetAttribute(probe_cm, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe_cm<<<1, 128, 32768>>>(tm, dout, plane * Hp + ph0, i0, pw0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> got(8192);
  cudaMemcpy(got.data(), dout, got.size() * 2, cudaMemcpyDeviceToHost);
  long bad = 0, zeros_ok = 0;
  for (int pw = 0; pw < 16; ++pw)
    for (int i = 0; i < 4; ++i)
      for (int jh = 0; jh < 2; ++jh)
        for (int r8 = 0; r8 < 8; ++r8)
          for (int j8 = 0; j8 < 8; ++j8) {
            const int idx = (((pw * 4 + i) * 2 + jh) * 8 + r8) * 8 + j8;
            if (pw >= Wp) {  // out of range along pw: zero fill
              if (got[idx] != 0) ++bad; else ++zeros_ok;
              continue;
            }
            const size_t Rr = size_t(plane) * Hp + ph0 + r8;  // rows past the plane continue into the next one
            const size_t src = Rr * 16 * W + size_t(i0 + i) * W + size_t(pw0 + pw) * 16 + jh * 8 + j8;
            if (src >= img.size() ? got[idx] != 0 : got[idx] != img[src]) ++bad;
          }
  printf("core-matrix layout [pw][i][jh][ph%%8][8]: %ld mismatches of 8192 (%ld zero-filled)\n", bad, zeros_ok);
  printf(bad ? "PROBE FAIL\n" : "PROBE PASS\n");
  return bad != 0;
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  if (variant == 3) return run_core_matrix_probe();  // 0: swizzle 128B, 1: no swizzle, 2: box of one patch row (ph box 1), 3: i box 16 (32-byte rows)

  const int N = 2, H = 224, W = 224, p = 16, Wp = W / p, Hp = H / p;
  std::vector<uint16_t> img(size_t(N) * 3 * H * W);
  for (size_t i = 0; i < img.size(); ++i) img[i] = uint16_t(i * 2654435761u >> 16);
  uint16_t *dimg, *dout;
  cudaMalloc(&dimg, img.size() * 2);
  cudaMalloc(&dout, 128 * 64 * 2);
  cudaMemcpy(dimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Fn fn = (Fn)fnp;
  CUtensorMap tm;
  cuuint64_t dims[5] = {cuuint64_t(p), cuuint64_t(p), cuuint64_t(Wp), cuuint64_t(Hp), cuuint64_t(N * 3)};
  cuuint64_t st[4] = {cuuint64_t(W) * 2, cuuint64_t(p) * 2, cuuint64_t(p) * W * 2, cuuint64_t(H) * W * 2};
  const int ph_box = variant == 2 ? 1 : 7;
  cuuint32_t box[5] = {cuuint32_t(p), 4, cuuint32_t(Wp), cuuint32_t(ph_box), 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dimg, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  variant == 1 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (j, i, pw, ph, nc) with strides {2W, 2p, 2pW, 2HW}: CUresult %d\n", int(r));
  if (r != CUDA_SUCCESS) return 1;
  const int rows = Wp * ph_box, bytes = rows * 128;
  const int i0 = 8, ph0 = 7, nc = 4;  // channel 1 of image 1, patch rows 7..13, pixel rows 8..11 of each patch
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(tm, dout, i0, ph0, nc, rows, bytes, variant == 1 ? 0 : 1);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint16_t> got(rows * 64);
  cudaMemcpy(got.data(), dout, got.size() * 2, cudaMemcpyDeviceToHost);
  long bad = 0;
  for (int rr = 0; rr < rows; ++rr)
    for (int e2 = 0; e2 < 64; ++e2) {
      const int ph = ph0 + rr / Wp, pw = rr % Wp, i = i0 + e2 / 16, j = e2 % 16;
      const size_t src = (size_t(nc) * H + ph * p + i) * W + pw * p + j;
      if (got[rr * 64 + e2] != img[src]) ++bad;
    }
  printf("box -> smem rows [patch][i][j]: %ld mismatches of %d\n", bad, rows * 64);
  printf(bad ? "PROBE FAIL\n" : "PROBE PASS\n");
  return bad != 0;
}
