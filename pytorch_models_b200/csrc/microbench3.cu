// Pipe microbenchmarks behind the short-sequence attention kernel: throughput and dependent-chain
// latency of FMNMX / FMNMX3 / FFMA2 / F2FP with 1, 2 and 4 warps per SM sub-partition.   make ../b200enc_microbench3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmax2v(float a, float b) {
  float d;
  asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// mode 0: FMNMX independent (8 chains), 1: FMNMX3 independent (8 chains), 2: FMNMX one dependent chain,
// 3: FMNMX3 one dependent chain, 4: F2FP pack independent
template <int kMode>
__global__ void k(const float* in, float* out, long long* clk, int iters) {
  float a[8], x = in[threadIdx.x], y = in[threadIdx.x + 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (kMode == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmax2v(a[i], x);
      } else if (kMode == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmax3(a[i], x, y);
      } else if (kMode == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[0] = fmax2v(a[0], a[(i & 3) + 1]);
      } else if (kMode == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[0] = fmax3(a[0], a[(i & 3) + 1], y);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 v = __floats2bfloat162_rn(a[i], x);
          uint32_t u = *reinterpret_cast<uint32_t*>(&v);
          asm volatile("" : "+r"(u));
          a[i] = __uint_as_float(u);
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int kMode>
static void run(const char* name) {
  float *in, *out;
  long long* clk;
  cudaMalloc(&in, 4096 * 4);
  cudaMemset(in, 0, 4096 * 4);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 8);
  const int iters = 2000;
  for (int threads : {128, 256, 512}) {
    k<kMode><<<148, threads>>>(in, out, clk, iters);
    k<kMode><<<148, threads>>>(in, out, clk, iters);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double n = double(iters) * 32;  // instructions per warp
    printf("%-34s %d warps/SMSP: %7.2f clk per warp-instruction, %6.2f clk per instr per SMSP\n", name, threads / 128,
           c / n, c / n / (threads / 128));
  }
}

int main() {
  run<0>("FMNMX, 8 independent chains");
  run<1>("FMNMX3, 8 independent chains");
  run<2>("FMNMX, one dependent chain");
  run<3>("FMNMX3, one dependent chain");
  run<4>("F2FP.BF16 pack, 8 independent");
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
