// C-ABI entry points for the HBM-bound row kernels: LayerNorm, row statistics, pooling, patch rows.
#include <algorithm>

#include "../../include/b200enc.h"
#include "host_util.h"
#include "layernorm.cuh"
#include "logmel.cuh"
#include "patchify.cuh"

using namespace b200;

static int check_rows(const char* who, const void* x, long long ldx, int rows, int d) {
  B200_CHECK_ARG(x != nullptr, "%s: null input", who);
  B200_CHECK_ARG(rows >= 1 && d >= 8 && d % 8 == 0 && d <= 8 * 32 * LN_MAX_CHUNKS,
                 "%s: rows=%d d=%d unsupported (d must be a multiple of 8, <= %d)", who, rows, d,
                 8 * 32 * LN_MAX_CHUNKS);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && ldx % 8 == 0,
                 "%s: rows must be 16-byte aligned", who);
  return 0;
}

// persistent grid: enough warps to fill every SM several times over, never more than one warp per row
template <bool kWriteOut>
static int launch_layernorm(const void* x, long long ldx, const float* gamma, const float* beta, float eps, int rows, int d,
                            void* out, long long ldo, float* stats, void* stream) {
  const int blocks_needed = (rows + LN_WARPS - 1) / LN_WARPS;
  const int grid = std::min(blocks_needed, sm_count() * 8);
  const int nc = (d / 8 + 31) / 32;
  auto go = [&](auto kern) {
    kern<<<grid, LN_WARPS * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, gamma, beta, eps, rows, d, reinterpret_cast<__nv_bfloat16*>(out), ldo,
        reinterpret_cast<float2*>(stats));
  };
  switch (nc) {
    case 1: go(layernorm_kernel<kWriteOut, 1>); break;
    case 2: go(layernorm_kernel<kWriteOut, 2>); break;
    case 3: go(layernorm_kernel<kWriteOut, 3>); break;
    case 4: go(layernorm_kernel<kWriteOut, 4>); break;
    case 5: go(layernorm_kernel<kWriteOut, 5>); break;
    case 6: go(layernorm_kernel<kWriteOut, 6>); break;
    case 7: go(layernorm_kernel<kWriteOut, 7>); break;
    default: go(layernorm_kernel<kWriteOut, 8>); break;
  }
  B200_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200enc_layernorm(const void* x, long long ldx, const float* gamma, const float* beta, float eps,
                                 int rows, int d, void* out, long long ldo, float* stats, void* stream) {
  int rc = check_rows("b200enc_layernorm", x, ldx, rows, d);
  if (rc) return rc;
  B200_CHECK_ARG(gamma && beta && out, "b200enc_layernorm: null pointer");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0 && ldo % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(gamma) & 15u) == 0 && (reinterpret_cast<uintptr_t>(beta) & 15u) == 0,
                 "b200enc_layernorm: out/gamma/beta must be 16-byte aligned");
  return launch_layernorm<true>(x, ldx, gamma, beta, eps, rows, d, out, ldo, stats, stream);
}

extern "C" int b200enc_row_stats(const void* x, long long ldx, float eps, int rows, int d, float* stats,
                                 void* stream) {
  int rc = check_rows("b200enc_row_stats", x, ldx, rows, d);
  if (rc) return rc;
  B200_CHECK_ARG(stats != nullptr, "b200enc_row_stats: null stats");
  return launch_layernorm<false>(x, ldx, nullptr, nullptr, eps, rows, d, nullptr, 0, stats, stream);
}

extern "C" int b200enc_mean_tokens(const void* x, long long batch_stride, long long ldx, int B, int L, int d,
                                   void* out, long long ldo, void* stream) {
  B200_CHECK_ARG(x && out && B >= 1 && L >= 1 && d >= 1, "b200enc_mean_tokens: bad arguments");
  dim3 grid((d + 255) / 256, B);
  mean_tokens_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), batch_stride, ldx, L, d, reinterpret_cast<__nv_bfloat16*>(out), ldo);
  B200_CUDA(cudaGetLastError());
  return 0;
}

template <typename TIn>
static int launch_patchify(const void* img, int B, int H, int W, int p, int Kpad, void* rows, cudaStream_t s) {
  const int V = (p % 8 == 0 && W % 8 == 0) ? 8 : ((p % 2 == 0 && W % 2 == 0) ? 2 : 1);
  const long long total = (long long)B * 3 * H * (W / V);
  const int grid = int((total + 255) / 256);
  const TIn* src = reinterpret_cast<const TIn*>(img);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(rows);
  if (V == 8)
    patchify_kernel<TIn, 8><<<grid, 256, 0, s>>>(src, B, H, W, p, Kpad, dst);
  else if (V == 2)
    patchify_kernel<TIn, 2><<<grid, 256, 0, s>>>(src, B, H, W, p, Kpad, dst);
  else
    patchify_kernel<TIn, 1><<<grid, 256, 0, s>>>(src, B, H, W, p, Kpad, dst);
  B200_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200enc_patch_rows(const void* img, int img_dtype, int B, int H, int W, int p, int Kpad, void* rows,
                                  void* stream) {
  B200_CHECK_ARG(img && rows, "b200enc_patch_rows: null pointer");
  B200_CHECK_ARG(B >= 1 && p >= 1 && H % p == 0 && W % p == 0, "b200enc_patch_rows: image %dx%d not divisible by patch %d",
                 H, W, p);
  B200_CHECK_ARG(Kpad >= 3 * p * p && Kpad % 8 == 0, "b200enc_patch_rows: Kpad=%d must be a multiple of 8 >= %d", Kpad,
                 3 * p * p);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15u) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15u) == 0,
                 "b200enc_patch_rows: pointers must be 16-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (img_dtype == B200ENC_DTYPE_BF16) return launch_patchify<__nv_bfloat16>(img, B, H, W, p, Kpad, rows, s);
  if (img_dtype == B200ENC_DTYPE_F32) return launch_patchify<float>(img, B, H, W, p, Kpad, rows, s);
  return set_error(-1, "b200enc_patch_rows: unsupported image dtype %d", img_dtype);
}

extern "C" int b200enc_cls_rows(const void* cls, int B, int d, void* tokens, long long batch_stride, void* stream) {
  B200_CHECK_ARG(cls && tokens && B >= 1 && d >= 1, "b200enc_cls_rows: bad arguments");
  const long long total = (long long)B * d;
  cls_rows_kernel<<<int((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(cls), B, d, reinterpret_cast<__nv_bfloat16*>(tokens), batch_stride);
  B200_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200enc_embed_rows(const long long* ids, long long rows, int L, const void* tok, const void* pos,
                                  int dtype, int vocab, int d, void* out, void* stream) {
  B200_CHECK_ARG(ids && tok && pos && out && rows >= 1 && L >= 1 && vocab >= 1, "b200enc_embed_rows: bad arguments");
  B200_CHECK_ARG(d >= 8 && d % 8 == 0, "b200enc_embed_rows: d=%d must be a multiple of 8", d);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "b200enc_embed_rows: out must be 16-byte aligned");
  const long long total = rows * (d / 8);
  const unsigned grid = unsigned((total + 255) / 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(out);
  if (dtype == B200ENC_DTYPE_BF16)
    embed_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(ids, rows, L, reinterpret_cast<const __nv_bfloat16*>(tok),
                                                          reinterpret_cast<const __nv_bfloat16*>(pos), vocab, d, dst);
  else if (dtype == B200ENC_DTYPE_F32)
    embed_rows_kernel<float><<<grid, 256, 0, s>>>(ids, rows, L, reinterpret_cast<const float*>(tok),
                                                  reinterpret_cast<const float*>(pos), vocab, d, dst);
  else
    return set_error(-1, "b200enc_embed_rows: unsupported dtype %d", dtype);
  B200_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200enc_time_rows(const void* x, int dtype, int N, int C, int T, void* rows, void* stream) {
  B200_CHECK_ARG(x && rows && N >= 1 && C >= 1 && T >= 1, "b200enc_time_rows: bad arguments");
  dim3 grid((T + 31) / 32, (C + 31) / 32, N);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(rows);
  if (dtype == B200ENC_DTYPE_BF16)
    time_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), C, T, dst);
  else if (dtype == B200ENC_DTYPE_F32)
    time_rows_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), C, T, dst);
  else
    return set_error(-1, "b200enc_time_rows: unsupported dtype %d", dtype);
  B200_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200enc_whisper_logmel(const float* audio, long long audio_stride, int N, int L, const float* filters_t,
                                      int n_mels, float* out, int* sample_max, void* stream) {
  B200_CHECK_ARG(audio && filters_t && out && sample_max, "b200enc_whisper_logmel: null pointer");
  B200_CHECK_ARG(N >= 1 && L > LM_NFFT / 2 && audio_stride >= L,
                 "b200enc_whisper_logmel: bad shape N=%d L=%d stride=%lld (reflect padding needs L > 200)", N, L,
                 audio_stride);
  B200_CHECK_ARG(n_mels >= 1 && n_mels <= LM_THREADS, "b200enc_whisper_logmel: n_mels=%d must be in [1, %d]", n_mels,
                 LM_THREADS);
  const int T = L / LM_HOP;  // torch.stft gives 1 + L / hop frames; the reference drops the last one
  B200_CHECK_ARG(T >= 1, "b200enc_whisper_logmel: fewer than %d samples", LM_HOP);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  lm_init_max_kernel<<<(N + 255) / 256, 256, 0, s>>>(sample_max, N);
  dim3 grid((T + LM_FRAMES - 1) / LM_FRAMES, N);
  logmel_kernel<<<grid, LM_THREADS, 0, s>>>(audio, audio_stride, L, T, filters_t, n_mels, out, sample_max);
  const long long per_sample = (long long)n_mels * T;
  dim3 grid2(unsigned(std::min<long long>((per_sample + 255) / 256, 1024)), N);
  logmel_finish_kernel<<<grid2, 256, 0, s>>>(out, per_sample, sample_max);
  B200_CUDA(cudaGetLastError());
  return 0;
}
