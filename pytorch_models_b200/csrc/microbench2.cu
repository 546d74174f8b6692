// Round-2 microbenchmarks behind the softmax restructuring of attention.cuh (not part of libb200enc.so):
// how long ONE warp (and two warps per SM sub-partition) need for the exponential phase and for the row-maximum phase
// of a 128-column block held in registers, for several source-level schedules; and the TMEM read rate per load shape.
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace b200;

__device__ __forceinline__ float ex2v(float x) {  // volatile: keeps its place relative to other volatile asm
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x = make_float2(fmaxf(x.x, -126.0f), fmaxf(x.y, -126.0f));
  const float2 magic = make_float2(12582912.0f, 12582912.0f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(n, make_float2(-1.0f, -1.0f), x);
  float2 q = __ffma2_rn(f, make_float2(0.05508868396282196f, 0.05508868396282196f),
                        make_float2(0.24260404706001282f, 0.24260404706001282f));
  q = __ffma2_rn(q, f, make_float2(0.6932762265205383f, 0.6932762265205383f));
  q = __ffma2_rn(q, f, make_float2(0.9999289512634277f, 0.9999289512634277f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}

// MODE 0: the shipped loop (interleaved, every 4th pair on the FMA pipe)      MODE 1: same without the polynomial
// MODE 2: two-stage pipeline, 8 pairs per stage (MUFU of stage k+1 before the sums/packs of stage k), volatile MUFU
// MODE 3: as 2 with 16 pairs per stage                                        MODE 4: as 2 + polynomial share
template <int MODE>
__global__ void exp_phase_kernel(float* out, long long* cycles, int iters) {
  float v[128];
#pragma unroll
  for (int k = 0; k < 128; ++k) v[k] = -(float)((threadIdx.x * 7 + k * 13) % 97) * 0.05f;
  uint32_t acc = 0;
  float2 sum0 = make_float2(0.f, 0.f), sum1 = sum0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float c = 0.18f + it * 1e-9f;
    const float2 c2 = make_float2(c, c), nm = make_float2(-0.01f, -0.01f);
    if (MODE == 0 || MODE == 1) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 e = __ffma2_rn(make_float2(v[ch * 32 + 2 * i], v[ch * 32 + 2 * i + 1]), c2, nm);
          const float2 pr = (MODE == 0 && (i % 4) == 3) ? exp2_poly2(e) : make_float2(fast_exp2(e.x), fast_exp2(e.y));
          if (i & 1) sum1 = __fadd2_rn(sum1, pr); else sum0 = __fadd2_rn(sum0, pr);
          pk[i] = pack_bf16x2(pr.x, pr.y);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= pk[i] + i;
      }
    } else {
      constexpr int S = MODE == 3 ? 16 : 8;      // pairs per stage
      constexpr int NS = 64 / S;
      float2 e[2][S];
      auto stage_mufu = [&](int s, int buf) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
          const int k = s * S + i;
          const float2 x = __ffma2_rn(make_float2(v[2 * k], v[2 * k + 1]), c2, nm);
          if (MODE == 4 && (i % 4) == 3) e[buf][i] = exp2_poly2(x);
          else e[buf][i] = make_float2(ex2v(x.x), ex2v(x.y));
        }
      };
      auto stage_use = [&](int buf) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
          if (i & 1) sum1 = __fadd2_rn(sum1, e[buf][i]); else sum0 = __fadd2_rn(sum0, e[buf][i]);
          acc ^= pack_bf16x2(e[buf][i].x, e[buf][i].y) + i;
        }
      };
      stage_mufu(0, 0);
#pragma unroll
      for (int s = 1; s < NS; ++s) {
        stage_mufu(s, s & 1);
        stage_use((s - 1) & 1);
      }
      stage_use((NS - 1) & 1);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum0.x + sum0.y + sum1.x + sum1.y + __uint_as_float(acc & 0x3f800000u);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Row maximum of 128 register values: MODE 0 two dependent chains (shipped), 1 eight chains + tree, 2 fmax3 tree
template <int MODE>
__global__ void max_phase_kernel(float* out, long long* cycles, int iters) {
  float v[128];
#pragma unroll
  for (int k = 0; k < 128; ++k) v[k] = (float)((threadIdx.x * 7 + k * 13) % 97) * 0.05f;
  float res = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float m;
    if (MODE == 0) {
      float a = -INFINITY, b = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 2) { a = fmaxf(a, v[i]); b = fmaxf(b, v[i + 1]); }
      m = fmaxf(a, b);
    } else if (MODE == 1) {
      float a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = v[j];
#pragma unroll
      for (int i = 8; i < 128; i += 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], v[i + j]);
      m = fmaxf(fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3])), fmaxf(fmaxf(a[4], a[5]), fmaxf(a[6], a[7])));
    } else {
      float a[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = fmaxf(fmaxf(v[8 * j], v[8 * j + 1]), fmaxf(fmaxf(v[8 * j + 2], v[8 * j + 3]), fmaxf(fmaxf(v[8 * j + 4], v[8 * j + 5]), fmaxf(v[8 * j + 6], v[8 * j + 7]))));
#pragma unroll
      for (int s = 8; s > 0; s >>= 1)
#pragma unroll
        for (int j = 0; j < s; ++j) a[j] = fmaxf(a[j], a[j + s]);
      m = a[0];
    }
    res += m;
    v[0] += 1e-3f * m;  // loop-carried: keeps the reduction inside the loop (constant index: v stays in registers)
    v[77] -= 1e-3f * m;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = res;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// TMEM read rate per load shape: SHAPE 0: 4 x (32x32b.x32), 1: 2 x (32x32b.x64), 2: 1 x (32x32b.x128)
template <int SHAPE>
__global__ void tmem_shape_kernel(float* out, long long* cycles, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc<512>(smem_u32(&slot));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (SHAPE == 0) {
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(base + c * 32, v[c]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += v[c][0] ^ v[c][13] ^ v[c][31];
    } else if (SHAPE == 1) {
      uint32_t v[2][64];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
            "%24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, "
            "%46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
            : "=r"(v[c][0]), "=r"(v[c][1]), "=r"(v[c][2]), "=r"(v[c][3]), "=r"(v[c][4]), "=r"(v[c][5]), "=r"(v[c][6]),
              "=r"(v[c][7]), "=r"(v[c][8]), "=r"(v[c][9]), "=r"(v[c][10]), "=r"(v[c][11]), "=r"(v[c][12]), "=r"(v[c][13]),
              "=r"(v[c][14]), "=r"(v[c][15]), "=r"(v[c][16]), "=r"(v[c][17]), "=r"(v[c][18]), "=r"(v[c][19]),
              "=r"(v[c][20]), "=r"(v[c][21]), "=r"(v[c][22]), "=r"(v[c][23]), "=r"(v[c][24]), "=r"(v[c][25]),
              "=r"(v[c][26]), "=r"(v[c][27]), "=r"(v[c][28]), "=r"(v[c][29]), "=r"(v[c][30]), "=r"(v[c][31]),
              "=r"(v[c][32]), "=r"(v[c][33]), "=r"(v[c][34]), "=r"(v[c][35]), "=r"(v[c][36]), "=r"(v[c][37]),
              "=r"(v[c][38]), "=r"(v[c][39]), "=r"(v[c][40]), "=r"(v[c][41]), "=r"(v[c][42]), "=r"(v[c][43]),
              "=r"(v[c][44]), "=r"(v[c][45]), "=r"(v[c][46]), "=r"(v[c][47]), "=r"(v[c][48]), "=r"(v[c][49]),
              "=r"(v[c][50]), "=r"(v[c][51]), "=r"(v[c][52]), "=r"(v[c][53]), "=r"(v[c][54]), "=r"(v[c][55]),
              "=r"(v[c][56]), "=r"(v[c][57]), "=r"(v[c][58]), "=r"(v[c][59]), "=r"(v[c][60]), "=r"(v[c][61]),
              "=r"(v[c][62]), "=r"(v[c][63])
            : "r"(base + c * 64)
            : "memory");
      }
      tmem_wait_ld();
      acc += v[0][0] ^ v[0][63] ^ v[1][5] ^ v[1][63];
    } else {
      // 16x256b shape: 16 lanes x 256 bits per repetition; x16 = 64 columns for the 32 lanes as two halves
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        asm volatile(
            "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
            "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[c][0]), "=r"(v[c][1]), "=r"(v[c][2]), "=r"(v[c][3]), "=r"(v[c][4]), "=r"(v[c][5]), "=r"(v[c][6]),
              "=r"(v[c][7]), "=r"(v[c][8]), "=r"(v[c][9]), "=r"(v[c][10]), "=r"(v[c][11]), "=r"(v[c][12]), "=r"(v[c][13]),
              "=r"(v[c][14]), "=r"(v[c][15]), "=r"(v[c][16]), "=r"(v[c][17]), "=r"(v[c][18]), "=r"(v[c][19]),
              "=r"(v[c][20]), "=r"(v[c][21]), "=r"(v[c][22]), "=r"(v[c][23]), "=r"(v[c][24]), "=r"(v[c][25]),
              "=r"(v[c][26]), "=r"(v[c][27]), "=r"(v[c][28]), "=r"(v[c][29]), "=r"(v[c][30]), "=r"(v[c][31])
            : "r"(base + (c & 1) * 64 + (uint32_t((c >> 1) * 16) << 16))  // two 16-lane halves x two 64-column groups
            : "memory");
      }
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += v[c][0] ^ v[c][13] ^ v[c][31];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(slot);
  }
}

// MUFU rate per operand type: MODE 0 ex2.approx.ftz.f32 (8 per iteration), 1 ex2.approx.f16x2 (8 packed = 16 values),
// 2 ex2.approx.ftz.bf16x2 (8 packed = 16 values)
template <int MODE>
__global__ void mufu_type_kernel(float* out, long long* cycles, int iters) {
  uint32_t h[8];
  float x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = -(threadIdx.x + k) * 1e-3f;
    h[k] = 0xB800B400u + threadIdx.x + k;  // two small negative halves
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[k]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[k]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k] + __uint_as_float(h[k] & 0x3f800000u);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename K>
static double run(K kern, int sms, int threads, int iters, float* out, long long* cyc) {
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<sms, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
  }
  long long h[256];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
  return double(mx) / iters;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 2000;
  const char* en[5] = {"shipped loop (25% poly)", "shipped loop, no poly", "2-stage pipeline, 8 pairs", "2-stage pipeline, 16 pairs",
                       "2-stage pipeline, 8 pairs, 25% poly"};
  for (int threads : {128, 256, 384}) {
    double r[5];
    r[0] = run(exp_phase_kernel<0>, sms, threads, iters, out, cyc);
    r[1] = run(exp_phase_kernel<1>, sms, threads, iters, out, cyc);
    r[2] = run(exp_phase_kernel<2>, sms, threads, iters, out, cyc);
    r[3] = run(exp_phase_kernel<3>, sms, threads, iters, out, cyc);
    r[4] = run(exp_phase_kernel<4>, sms, threads, iters, out, cyc);
    for (int m = 0; m < 5; ++m)
      printf("exp phase, 128 columns per thread, %-38s %d warps/SMSP: %7.1f clk per block\n", en[m], threads / 128, r[m]);
  }
  const char* mn[3] = {"2 dependent chains (shipped)", "8 chains + tree", "per-8 groups + tree"};
  for (int threads : {128, 256}) {
    double r[3];
    r[0] = run(max_phase_kernel<0>, sms, threads, iters, out, cyc);
    r[1] = run(max_phase_kernel<1>, sms, threads, iters, out, cyc);
    r[2] = run(max_phase_kernel<2>, sms, threads, iters, out, cyc);
    for (int m = 0; m < 3; ++m)
      printf("row max of 128 values, %-30s %d warps/SMSP: %7.1f clk per block\n", mn[m], threads / 128, r[m]);
  }
  const char* tn[3] = {"4 x 32x32b.x32", "2 x 32x32b.x64", "4 x 16x256b.x8"};
  for (int threads : {128, 256}) {
    double r[3];
    r[0] = run(tmem_shape_kernel<0>, sms, threads, iters, out, cyc);
    r[1] = run(tmem_shape_kernel<1>, sms, threads, iters, out, cyc);
    r[2] = run(tmem_shape_kernel<2>, sms, threads, iters, out, cyc);
    for (int m = 0; m < 3; ++m)
      printf("TMEM read of 128 columns x 32 lanes per warp, %-16s %d warps/SMSP: %7.1f clk per 16 KB -> %.1f B/clk/SM\n", tn[m],
             threads / 128, r[m], double(threads / 32) * 16384.0 / r[m]);
  }
  for (int threads : {128, 512}) {
    const double a = run(mufu_type_kernel<0>, sms, threads, iters, out, cyc);
    const double b = run(mufu_type_kernel<1>, sms, threads, iters, out, cyc);
    const double c = run(mufu_type_kernel<2>, sms, threads, iters, out, cyc);
    printf("MUFU.EX2 values per clk per SM, %d warps/SMSP: f32 %.1f   f16x2 %.1f   bf16x2 %.1f\n", threads / 128,
           threads * 8.0 / a, threads * 16.0 / b, threads * 16.0 / c);
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
