// HBM-bound row kernels of the encoder path (reference: nn.LayerNorm at transformer.py:87,93, vit.py:69, whisper.py:27,
// bert.py:31 -> aten.native_layer_norm: biased variance, eps inside the sqrt).
//   row_stats_kernel : (mean, rstd) per row, consumed by the LayerNorm-folded GEMM epilogue
//   layernorm_kernel : full LayerNorm, bf16 in / bf16 out, arbitrary row strides (so it can gather e.g. only the
//                      class-token rows for the final norm + pooling)
// One warp per row, 16-byte loads, the row is held in registers so the variance is the exact two-pass form.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int LN_WARPS = 8;
constexpr int LN_MAX_CHUNKS = 8;  // per lane: d <= 8 * 32 * 8 = 2048

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// kNC = 16-byte chunks per lane (ceil(d / 256)): compile-time, so every load of a row is issued back to back and the
// row lives in kNC x 8 registers. Persistent: a warp walks rows with a grid stride, keeps gamma / beta in registers
// (round 1 re-read them from L1 for every row: four times the bytes of the row itself) and has the NEXT row's loads in
// flight while it reduces and writes the current one. 201 728 x 768: 197 -> see DESIGN.md section 3.3.
template <bool kWriteOut, int kNC>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, int rows, int d, __nv_bfloat16* __restrict__ out,
                 long long ldo, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * LN_WARPS;
  int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nchunks = d >> 3;  // 16-byte chunks in the row
  const float inv_d = 1.0f / float(d);
  float4 g[kWriteOut ? kNC : 1][2], bt[kWriteOut ? kNC : 1][2];
  if (kWriteOut) {
#pragma unroll
    for (int i = 0; i < kNC; ++i) {
      const int ch = lane + 32 * i;
      if (ch < nchunks) {
        g[i][0] = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch);
        g[i][1] = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * ch + 1);
        bt[i][0] = __ldg(reinterpret_cast<const float4*>(beta) + 2 * ch);
        bt[i][1] = __ldg(reinterpret_cast<const float4*>(beta) + 2 * ch + 1);
      }
    }
  }
  uint4 nxt[kNC];
  auto load_row = [&](int r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)r * ldx);
#pragma unroll
    for (int i = 0; i < kNC; ++i) {
      const int ch = lane + 32 * i;
      nxt[i] = make_uint4(0, 0, 0, 0);
      if (ch < nchunks) nxt[i] = __ldg(xr + ch);
    }
  };
  load_row(row);
  while (row < rows) {
    float v[kNC][8];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kNC; ++i) {
      const uint32_t w[4] = {nxt[i].x, nxt[i].y, nxt[i].z, nxt[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[i][2 * k] = bf16_lo(w[k]);
        v[i][2 * k + 1] = bf16_hi(w[k]);
        s += v[i][2 * k] + v[i][2 * k + 1];  // chunks past the row's end are zeros
      }
    }
    const int next = row + warps_total;
    if (next < rows) load_row(next);  // in flight during the reductions and the stores below
    const float mean = warp_sum(s) * inv_d;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < kNC; ++i) {
      if (lane + 32 * i < nchunks) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float c = v[i][k] - mean;
          ss = fmaf(c, c, ss);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_d + eps);
    if (stats != nullptr && lane == 0) stats[row] = make_float2(mean, rstd);
    if (kWriteOut) {
      uint4* orow = reinterpret_cast<uint4*>(out + (long long)row * ldo);
#pragma unroll
      for (int i = 0; i < kNC; ++i) {
        const int ch = lane + 32 * i;
        if (ch < nchunks) {
          const float4 g0 = g[i][0], g1 = g[i][1], b0 = bt[i][0], b1 = bt[i][1];
          uint4 w;
          w.x = pack_bf16x2(fmaf((v[i][0] - mean) * rstd, g0.x, b0.x), fmaf((v[i][1] - mean) * rstd, g0.y, b0.y));
          w.y = pack_bf16x2(fmaf((v[i][2] - mean) * rstd, g0.z, b0.z), fmaf((v[i][3] - mean) * rstd, g0.w, b0.w));
          w.z = pack_bf16x2(fmaf((v[i][4] - mean) * rstd, g1.x, b1.x), fmaf((v[i][5] - mean) * rstd, g1.y, b1.y));
          w.w = pack_bf16x2(fmaf((v[i][6] - mean) * rstd, g1.z, b1.z), fmaf((v[i][7] - mean) * rstd, g1.w, b1.w));
          orow[ch] = w;
        }
      }
    }
    row = next;
  }
}

// Mean over the token axis: out[b][c] = mean_l x[b][l][c]  (GlobalAveragePooling, vit.py:25-27).
__global__ void __launch_bounds__(256)
mean_tokens_kernel(const __nv_bfloat16* __restrict__ x, long long batch_stride, long long ldx, int L, int d,
                   __nv_bfloat16* __restrict__ out, long long ldo) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const __nv_bfloat16* p = x + (long long)b * batch_stride + c;
  float s = 0.0f;
  for (int l = 0; l < L; ++l) s += __bfloat162float(p[(long long)l * ldx]);
  out[(long long)b * ldo + c] = __float2bfloat16_rn(s / float(L));
}

}  // namespace b200
