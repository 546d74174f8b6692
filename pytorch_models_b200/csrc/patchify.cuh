// Patch rows for the patch-embedding GEMM (reference: nn.Conv2d(3, d, p, p) at image/vit.py:64,78, stride == kernel,
// followed by flatten(-2).transpose(-1,-2)): rows[b*P + ph*Wp + pw][c*p*p + i*p + j] = img[b][c][ph*p+i][pw*p+j],
// columns [3*p*p, Kpad) zero-filled so the row stride is a multiple of 16 bytes (p = 14: 588 -> 592).
// Also writes the class-token row of the token matrix (vit.py:80-81) when asked to.
#pragma once
#include "ptx.cuh"

namespace b200 {

template <typename TIn, int V>
__global__ void __launch_bounds__(256)
patchify_kernel(const TIn* __restrict__ img, int B, int H, int W, int p, int Kpad, __nv_bfloat16* __restrict__ rows) {
  const int Wp = W / p, Hp = H / p;
  const int wv = W / V;  // vector chunks per image row
  const long long total = (long long)B * 3 * H * wv;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) {
    const int xv = int(t % wv);
    long long r = t / wv;
    const int y = int(r % H);
    r /= H;
    const int c = int(r % 3);
    const int b = int(r / 3);
    const int x = xv * V;
    const int ph = y / p, i = y - ph * p, pw = x / p, j = x - pw * p;
    const TIn* src = img + (((long long)b * 3 + c) * H + y) * W + x;
    __nv_bfloat16* dst = rows + ((long long)b * Hp * Wp + (long long)ph * Wp + pw) * Kpad + (c * p + i) * p + j;
    __nv_bfloat16 tmp[V];
#pragma unroll
    for (int k = 0; k < V; ++k) tmp[k] = __float2bfloat16_rn(float(src[k]));
    if (V == 8) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(tmp);
    } else if (V == 2) {
      *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(tmp);
    } else {
#pragma unroll
      for (int k = 0; k < V; ++k) dst[k] = tmp[k];
    }
  }
  // zero the padding columns
  const int K = 3 * p * p;
  const int pad = Kpad - K;
  if (pad > 0) {
    const long long nrows = (long long)B * Hp * Wp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = t; q < nrows * pad; q += stride) rows[(q / pad) * Kpad + K + (q % pad)] = __float2bfloat16_rn(0.f);
  }
}

// tokens[b][0][:] = cls[:]  (class token gets no positional embedding: vit.py:79-81)
__global__ void __launch_bounds__(256)
cls_rows_kernel(const __nv_bfloat16* __restrict__ cls, int B, int d, __nv_bfloat16* __restrict__ tokens,
                long long batch_stride) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * d) return;
  const int b = int(t / d), c = int(t % d);
  tokens[(long long)b * batch_stride + c] = cls[c];
}

// Token + position embedding (text/bert.py:35-36, text/gpt2.py:22-23, audio2text/whisper.py:47-48):
// out[r][:] = bf16(tok[ids[r]][:] + pos[r % L][:]); 8 columns per thread. An id outside [0, vocab) cannot raise from
// device code, so its row is filled with NaN instead of reading out of bounds.
template <typename TIn>
__global__ void __launch_bounds__(256)
embed_rows_kernel(const long long* __restrict__ ids, long long rows, int L, const TIn* __restrict__ tok,
                  const TIn* __restrict__ pos, int vocab, int d, __nv_bfloat16* __restrict__ out) {
  const int vec_per_row = d >> 3;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * vec_per_row) return;
  const long long r = t / vec_per_row;
  const int c = int(t % vec_per_row) * 8;
  const long long id = ids[r];
  const bool ok = id >= 0 && id < vocab;
  const TIn* trow = tok + (ok ? id : 0) * (long long)d + c;
  const TIn* prow = pos + (long long)(r % L) * d + c;
  uint32_t packed[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = ok ? float(trow[2 * i]) + float(prow[2 * i]) : __int_as_float(0x7fc00000);
    const float b = ok ? float(trow[2 * i + 1]) + float(prow[2 * i + 1]) : __int_as_float(0x7fc00000);
    packed[i] = pack_bf16x2(a, b);
  }
  *reinterpret_cast<uint4*>(out + r * d + c) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
}

// Whisper stem input (audio2text/whisper.py:16-21,30): x (N, C, T) channel-major -> rows (N, T+2, C) time-major bf16,
// rows 0 and T+1 zero. A k=3, pad=1 Conv1d over time then reads, for output step t, the 3*C contiguous values that
// start at row t (stride 1) or 2t (stride 2): the convolution becomes a plain GEMM over an overlapping strided view.
template <typename TIn>
__global__ void __launch_bounds__(256)
time_rows_kernel(const TIn* __restrict__ x, int C, int T, __nv_bfloat16* __restrict__ rows) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const TIn* xn = x + (long long)n * C * T;
  __nv_bfloat16* rn = rows + (long long)n * (T + 2) * C;
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, t = t0 + tx;
    tile[i][tx] = (c < C && t < T) ? float(xn[(long long)c * T + t]) : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, c = c0 + tx;
    if (t < T && c < C) rn[(long long)(t + 1) * C + c] = __float2bfloat16_rn(tile[tx][i]);
  }
  if (blockIdx.x == 0 && ty == 0) {
    const int c = c0 + tx;
    if (c < C) {
      rn[c] = __float2bfloat16_rn(0.f);
      rn[(long long)(T + 1) * C + c] = __float2bfloat16_rn(0.f);
    }
  }
}

}  // namespace b200
