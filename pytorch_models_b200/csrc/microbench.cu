// Stand-alone pipe-throughput microbenchmarks used to size the attention softmax (not part of libb200enc.so):
// MUFU.EX2 rate, packed-FP32 (FFMA2) rate and the mix the softmax runs, per SM and per clock.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int MODE>
__global__ void pipe_kernel(float* out, long long* cycles, int iters) {
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = (threadIdx.x + k) * 1e-3f;
  float2 a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) a[k] = make_float2(x[2 * k], x[2 * k + 1]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {  // 8 independent MUFU.EX2 per iteration
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = ex2(x[k]);
    } else if (MODE == 1) {  // 8 FFMA2 per iteration (16 fp32 FMA)
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] = __ffma2_rn(a[k], make_float2(0.999f, 1.001f), make_float2(1e-3f, -1e-3f));
    } else {  // softmax mix per 2 elements: 1 FFMA2, 2 EX2, 1 FADD2, 1 pack
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 e = __ffma2_rn(a[k], make_float2(0.5f, 0.5f), make_float2(-1.0f, -1.0f));
        e = make_float2(ex2(e.x), ex2(e.y));
        a[k] = __fadd2_rn(a[k], e);
        __nv_bfloat162 pk = __floats2bfloat162_rn(e.x, e.y);
        x[k] += __uint_as_float(*reinterpret_cast<unsigned*>(&pk) & 0x3f800000u);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
#pragma unroll
  for (int k = 0; k < 4; ++k) s += a[k].x + a[k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

#include <cuda_bf16.h>

// mode 3/4: the real softmax inner loop on a 64-value row slice held in registers. mode 3 = compiler's schedule
// (interleaved FFMA2 / EX2 / FADD2 / pack), mode 4 = batched: all FFMA2, then all EX2 back to back, then sums + packs.
#define BARRIER16(a, o)                                                                                      \
  asm volatile("" : "+f"(a[o + 0]), "+f"(a[o + 1]), "+f"(a[o + 2]), "+f"(a[o + 3]), "+f"(a[o + 4]),          \
               "+f"(a[o + 5]), "+f"(a[o + 6]), "+f"(a[o + 7]), "+f"(a[o + 8]), "+f"(a[o + 9]), "+f"(a[o + 10]), \
               "+f"(a[o + 11]), "+f"(a[o + 12]), "+f"(a[o + 13]), "+f"(a[o + 14]), "+f"(a[o + 15]))
template <int MODE>
__global__ void softmax_kernel(float* out, long long* cycles, int iters) {
  float v[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) v[k] = (threadIdx.x + k) * 1e-3f;
  float2 sum0 = make_float2(0.f, 0.f), sum1 = sum0;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float2 c2 = make_float2(0.5f, 0.5f), m2 = make_float2(-1.0f - it * 1e-9f, -1.0f);
    float e[64];
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float2 t = __ffma2_rn(make_float2(v[2 * i], v[2 * i + 1]), c2, m2);
        const float2 pr = make_float2(ex2(t.x), ex2(t.y));
        if (i & 1) sum1 = __fadd2_rn(sum1, pr); else sum0 = __fadd2_rn(sum0, pr);
        __nv_bfloat162 pk = __floats2bfloat162_rn(pr.x, pr.y);
        acc ^= *reinterpret_cast<unsigned*>(&pk);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float2 t = __ffma2_rn(make_float2(v[2 * i], v[2 * i + 1]), c2, m2);
        e[2 * i] = t.x;
        e[2 * i + 1] = t.y;
      }
      BARRIER16(e, 0); BARRIER16(e, 16); BARRIER16(e, 32); BARRIER16(e, 48);
#pragma unroll
      for (int i = 0; i < 64; ++i) e[i] = ex2(e[i]);
      BARRIER16(e, 0); BARRIER16(e, 16); BARRIER16(e, 32); BARRIER16(e, 48);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float2 pr = make_float2(e[2 * i], e[2 * i + 1]);
        if (i & 1) sum1 = __fadd2_rn(sum1, pr); else sum0 = __fadd2_rn(sum0, pr);
        __nv_bfloat162 pk = __floats2bfloat162_rn(pr.x, pr.y);
        acc ^= *reinterpret_cast<unsigned*>(&pk);
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum0.x + sum0.y + sum1.x + sum1.y + __uint_as_float(acc & 0x3f800000u);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// mode 5: which companion instruction breaks the MUFU cadence of a lone warp? FLAGS bit0 FFMA2, bit1 FADD2,
// bit2 F2FP pack, bit3 PRMT (truncating) pack, bit4 plain FADD instead of FADD2
template <int FLAGS>
__global__ void mix_kernel(float* out, long long* cycles, int iters) {
  float v[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) v[k] = (threadIdx.x + k) * 1e-3f;
  float2 sum0 = make_float2(0.f, 0.f), sum1 = sum0;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float2 c2 = make_float2(0.5f, 0.5f), m2 = make_float2(-1.0f - it * 1e-9f, -1.0f);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float2 t = make_float2(v[2 * i], v[2 * i + 1]);
      if (FLAGS & 1) t = __ffma2_rn(t, c2, m2);
      const float2 pr = make_float2(ex2(t.x), ex2(t.y));
      if (FLAGS & 2) { if (i & 1) sum1 = __fadd2_rn(sum1, pr); else sum0 = __fadd2_rn(sum0, pr); }
      if (FLAGS & 16) { sum0.x += pr.x; sum1.x += pr.y; }
      if (FLAGS & 4) {
        __nv_bfloat162 pk = __floats2bfloat162_rn(pr.x, pr.y);
        acc ^= *reinterpret_cast<unsigned*>(&pk);
      }
      if (FLAGS & 8) acc ^= __byte_perm(__float_as_uint(pr.x), __float_as_uint(pr.y), 0x7632);
      if (!(FLAGS & 30)) acc ^= __float_as_uint(pr.x) ^ __float_as_uint(pr.y);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum0.x + sum0.y + sum1.x + sum1.y + __uint_as_float(acc & 0x3f800000u);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int FLAGS>
void run_mix(const char* name, int sms, float* out, long long* cyc) {
  for (int threads : {128, 256}) {
    const int it2 = 1024;
    for (int rep = 0; rep < 2; ++rep) {
      mix_kernel<FLAGS><<<sms, threads>>>(out, cyc, it2);
      cudaDeviceSynchronize();
    }
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mix %-34s %4d threads/SM: %.2f exp per clk per SM\n", name, threads, double(threads) * it2 * 64 / double(mx));
  }
}

// TMEM read / write throughput: every warp of the CTA streams its 32-lane quarter with tcgen05.ld / tcgen05.st x32.
#include "ptx.cuh"
template <bool kStore>
__global__ void tmem_kernel(float* out, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    b200::tmem_alloc<512>(b200::smem_u32(&slot));
    b200::tmem_relinquish();
  }
  b200::tc_fence_before();
  __syncthreads();
  b200::tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t col = ((warp >> 2) * 128 + c * 32) & 511;
      if (kStore) {
        b200::tmem_st32(base + col, v);
      } else {
        b200::tmem_ld32(base + col, v);
      }
    }
    if (kStore) b200::tmem_wait_st(); else b200::tmem_wait_ld();
    acc += v[it & 31];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  b200::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    b200::tc_fence_after();
    b200::tmem_dealloc<512>(slot);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 4096;
  const char* names[3] = {"MUFU.EX2 (8 per iter)", "FFMA2 (8 per iter)", "softmax mix (8 elements per iter)"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) pipe_kernel<0><<<sms, threads>>>(out, cyc, iters);
        if (mode == 1) pipe_kernel<1><<<sms, threads>>>(out, cyc, iters);
        if (mode == 2) pipe_kernel<2><<<sms, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[256];
      cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per_clk = double(threads) * iters * 8 / double(mx);
      printf("%-36s %4d threads/SM: %8lld clk -> %.2f per clk per SM\n", names[mode], threads, mx, per_clk);
    }
  }
  for (int mode = 3; mode <= 4; ++mode) {
    for (int threads : {128, 256, 384, 512}) {
      const int it2 = 1024;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 3) softmax_kernel<3><<<sms, threads>>>(out, cyc, it2);
        if (mode == 4) softmax_kernel<4><<<sms, threads>>>(out, cyc, it2);
        cudaDeviceSynchronize();
      }
      long long h[256];
      cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-36s %4d threads/SM: %8lld clk -> %.2f exp per clk per SM\n",
             mode == 3 ? "softmax row slice, compiler schedule" : "softmax row slice, batched EX2", threads, mx,
             double(threads) * it2 * 64 / double(mx));
    }
  }
  for (int store = 0; store < 2; ++store) {
    for (int threads : {128, 256, 512}) {
      const int it3 = 2048;
      for (int rep = 0; rep < 2; ++rep) {
        if (store) tmem_kernel<true><<<sms, threads>>>(out, cyc, it3); else tmem_kernel<false><<<sms, threads>>>(out, cyc, it3);
        cudaDeviceSynchronize();
      }
      long long h[256];
      cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("TMEM %s x32, %3d threads/SM: %.1f bytes per clk per SM\n", store ? "st" : "ld", threads,
             double(threads) * it3 * 4 * 32 * 4 / double(mx));
    }
  }
  run_mix<0>("EX2 + LOP", sms, out, cyc);
  run_mix<1>("FFMA2 + EX2 + LOP", sms, out, cyc);
  run_mix<3>("FFMA2 + EX2 + FADD2", sms, out, cyc);
  run_mix<17>("FFMA2 + EX2 + 2 FADD", sms, out, cyc);
  run_mix<5>("FFMA2 + EX2 + F2FP", sms, out, cyc);
  run_mix<9>("FFMA2 + EX2 + PRMT", sms, out, cyc);
  run_mix<7>("FFMA2 + EX2 + FADD2 + F2FP", sms, out, cyc);
  run_mix<11>("FFMA2 + EX2 + FADD2 + PRMT", sms, out, cyc);
  run_mix<25>("FFMA2 + EX2 + 2 FADD + PRMT", sms, out, cyc);
  {
    int o_plain = 0, o_tmem = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o_plain, pipe_kernel<0>, 256, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o_tmem, tmem_kernel<false>, 256, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, tmem_kernel<false>);
    printf("occupancy (CTAs of 256 threads per SM): plain kernel %d, kernel that uses tcgen05.alloc %d (regs %d)\n", o_plain,
           o_tmem, fa.numRegs);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
