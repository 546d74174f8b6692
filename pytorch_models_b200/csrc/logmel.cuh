// Whisper audio front end (reference: WhisperPreprocessor / MelSpectrogram / Spectrogram,
// pytorch_models/audio2text/whisper.py:138-148, pytorch_models/audio/spectrogram.py:7-45):
//   torch.stft(x, 400, 160, window=hann(400), center=True, pad_mode="reflect").abs()**2  ->  filters @ power
//   -> drop the last frame -> clamp(0).log10() -> max(x, max_over_sample - 8) -> (x + 4) / 4
// fp32 throughout (this is not tensor-core work: 1 GFLOP per 30 s sample as a direct 400-point DFT, ~1 % of the
// encoder that consumes it). One CTA = 8 consecutive frames of one sample:
//   phase 1  windowed samples of the 8 frames -> smem (reflect padding by index arithmetic)
//   phase 2  thread k = frequency bin k (201 bins): direct DFT with a 400-entry twiddle table in smem, 8 frames at once
//   phase 3  thread m = mel band m: sum_k filters[m][k] * power[k]; log10; per-sample maximum via an ordered-int atomic
// A second tiny kernel applies the per-sample dynamic-range floor and the affine rescale.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int LM_NFFT = 400;
constexpr int LM_HOP = 160;
constexpr int LM_BINS = LM_NFFT / 2 + 1;  // 201
constexpr int LM_FRAMES = 8;              // frames per CTA
constexpr int LM_THREADS = 256;

// order-preserving float <-> int mapping so that atomicMax works on negative values and -inf
__device__ __forceinline__ int lm_float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float lm_ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void lm_init_max_kernel(int* sample_max, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sample_max[i] = lm_float_to_ordered(-INFINITY);
}

// audio: [N][audio_stride] fp32 (L valid samples); filters_t: [201][n_mels] (transposed mel filter bank);
// out: [N][n_mels][T] with T = L / 160, holding log10(mel power) after this kernel; sample_max: [N] ordered ints
__global__ void __launch_bounds__(LM_THREADS)
logmel_kernel(const float* __restrict__ audio, long long audio_stride, int L, int T, const float* __restrict__ filters_t,
              int n_mels, float* __restrict__ out, int* __restrict__ sample_max) {
  __shared__ float xw[LM_FRAMES][LM_NFFT];       // windowed samples
  __shared__ float tw_c[LM_NFFT], tw_s[LM_NFFT]; // cos / sin of 2 pi i / 400
  __shared__ float pw[LM_FRAMES][LM_BINS + 7];   // power spectrum
  __shared__ float red[LM_THREADS / 32];
  const int n = blockIdx.y;
  const int t0 = blockIdx.x * LM_FRAMES;
  const float* x = audio + (long long)n * audio_stride;
  for (int i = threadIdx.x; i < LM_NFFT; i += LM_THREADS) {
    float s, c;
    sincospif(2.0f * float(i) / float(LM_NFFT), &s, &c);
    tw_c[i] = c;
    tw_s[i] = s;
  }
  for (int i = threadIdx.x; i < LM_FRAMES * LM_NFFT; i += LM_THREADS) {
    const int f = i / LM_NFFT, j = i % LM_NFFT;
    // frame t covers padded samples [160 t, 160 t + 400), padded[p] = x[reflect(p - 200)]  (torch.stft center=True)
    int idx = (t0 + f) * LM_HOP + j - LM_NFFT / 2;
    if (idx < 0) idx = -idx;
    if (idx >= L) idx = 2 * (L - 1) - idx;
    idx = max(0, min(L - 1, idx));  // only frames beyond T (never stored) can still be out of range
    const float w = 0.5f - 0.5f * cospif(2.0f * float(j) / float(LM_NFFT));  // periodic Hann window
    xw[f][j] = w * x[idx];
  }
  __syncthreads();
  const int k = threadIdx.x;
  if (k < LM_BINS) {
    float re[LM_FRAMES], im[LM_FRAMES];
#pragma unroll
    for (int f = 0; f < LM_FRAMES; ++f) re[f] = im[f] = 0.0f;
    int ph = 0;  // (k * j) mod 400
    for (int j = 0; j < LM_NFFT; ++j) {
      const float c = tw_c[ph], s = tw_s[ph];
#pragma unroll
      for (int f = 0; f < LM_FRAMES; ++f) {
        const float v = xw[f][j];
        re[f] = fmaf(v, c, re[f]);
        im[f] = fmaf(v, s, im[f]);
      }
      ph += k;
      if (ph >= LM_NFFT) ph -= LM_NFFT;
    }
#pragma unroll
    for (int f = 0; f < LM_FRAMES; ++f) pw[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();
  float local_max = -INFINITY;
  const int m = threadIdx.x;
  if (m < n_mels) {
    float acc[LM_FRAMES];
#pragma unroll
    for (int f = 0; f < LM_FRAMES; ++f) acc[f] = 0.0f;
    for (int kk = 0; kk < LM_BINS; ++kk) {
      const float w = filters_t[kk * n_mels + m];
#pragma unroll
      for (int f = 0; f < LM_FRAMES; ++f) acc[f] = fmaf(w, pw[f][kk], acc[f]);
    }
    float* orow = out + ((long long)n * n_mels + m) * T;
#pragma unroll
    for (int f = 0; f < LM_FRAMES; ++f) {
      if (t0 + f < T) {
        const float v = log10f(fmaxf(acc[f], 0.0f));
        orow[t0 + f] = v;
        local_max = fmaxf(local_max, v);
      }
    }
  }
  // per-sample maximum: warp shuffle, then one atomic per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local_max;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = red[0];
    for (int w = 1; w < LM_THREADS / 32; ++w) mx = fmaxf(mx, red[w]);
    atomicMax(sample_max + n, lm_float_to_ordered(mx));
  }
}

// x = (max(x, sample_max - 8) + 4) / 4, in place over [N][per_sample]
__global__ void logmel_finish_kernel(float* __restrict__ out, long long per_sample, const int* __restrict__ sample_max) {
  const int n = blockIdx.y;
  const float floor_v = lm_ordered_to_float(sample_max[n]) - 8.0f;
  float* o = out + (long long)n * per_sample;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x)
    o[i] = (fmaxf(o[i], floor_v) + 4.0f) * 0.25f;
}

}  // namespace b200
