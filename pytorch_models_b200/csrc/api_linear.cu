// C-ABI entry point for the linear layers of the encoder block (see include/b200enc.h).
#include "../../include/b200enc.h"
#include "gemm.cuh"
#include "host_util.h"

using namespace b200;

namespace {

template <bool kFold, bool kGelu, bool kRes, bool kTma>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p, int grid,
                cudaStream_t stream) {
  auto kern = gemm_bf16_kernel<kFold, kGelu, kRes, kTma>;
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ta, tb, tc, p);
  B200_CUDA(cudaGetLastError());
  return 0;
}

template <bool kFold, bool kGelu, bool kRes>
int dispatch_store(bool tma, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const GemmParams& p,
                   int grid, cudaStream_t s) {
  return tma ? launch_gemm<kFold, kGelu, kRes, true>(ta, tb, tc, p, grid, s)
             : launch_gemm<kFold, kGelu, kRes, false>(ta, tb, tc, p, grid, s);
}

}  // namespace

extern "C" int b200enc_linear(const void* x, long long x_batch_stride, int ldx, const void* w, int ldw,
                              const float* bias, const float* colsum, const float* rowstats, const void* residual,
                              long long res_batch_stride, int ldr, void* out, long long out_batch_stride, int ldo,
                              int batches, int M, int N, int K, int flags, void* stream) {
  B200_CHECK_ARG(x && w && out, "b200enc_linear: null tensor pointer");
  B200_CHECK_ARG(batches >= 1 && M >= 1 && N >= 1 && K >= 1, "b200enc_linear: bad shape batches=%d M=%d N=%d K=%d",
                 batches, M, N, K);
  B200_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "b200enc_linear: K=%d and N=%d must be multiples of 8", K, N);
  // ldx < K is allowed on purpose: overlapping rows express a strided 1-D convolution as a GEMM (whisper stem).
  B200_CHECK_ARG(ldx >= 8 && ldw >= K && ldo >= N, "b200enc_linear: leading dimension smaller than the row length");
  B200_CHECK_ARG((colsum == nullptr) == (rowstats == nullptr),
                 "b200enc_linear: colsum and rowstats must be given together (LayerNorm fold)");
  if (residual) {
    B200_CHECK_ARG(ldr >= N && ldr % 8 == 0 && res_batch_stride % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(residual) & 15u) == 0,
                   "b200enc_linear: residual must be 16-byte aligned with strides that are multiples of 8");
  }
  const bool fold = colsum != nullptr;
  const bool gelu = (flags & B200ENC_LINEAR_GELU) != 0;
  const bool res = residual != nullptr;
  const bool tma_store = (flags & B200ENC_LINEAR_DIRECT_STORE) == 0;

  CUtensorMap ta, tb, tc;
  int rc;
  if ((rc = make_tmap_bf16(&ta, x, K, M, batches, ldx, batches > 1 ? x_batch_stride : (long long)M * ldx, GEMM_BK,
                           GEMM_BM, 128)))
    return rc;
  if ((rc = make_tmap_bf16(&tb, w, K, N, 0, ldw, 0, GEMM_BK, GEMM_BN, 128))) return rc;
  if ((rc = make_tmap_bf16(&tc, out, N, M, batches, ldo, batches > 1 ? out_batch_stride : (long long)M * ldo, 64, 32,
                           128)))
    return rc;
  if (!tma_store) {
    B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0 && ldo % 8 == 0 && out_batch_stride % 8 == 0,
                   "b200enc_linear: direct-store path needs 16-byte aligned rows");
  }

  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.batches = batches;
  p.tiles_m = (M + GEMM_BM - 1) / GEMM_BM;
  p.tiles_n = (N + GEMM_BN - 1) / GEMM_BN;
  p.bias = bias;
  p.colsum = colsum;
  p.rowstats = reinterpret_cast<const float2*>(rowstats);
  p.res = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.res_batch_stride = res_batch_stride;
  p.ldr = ldr;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.out_batch_stride = out_batch_stride;
  p.ldo = ldo;

  const long long total = (long long)p.tiles_m * p.tiles_n * batches;
  const int grid = int(total < sm_count() ? total : sm_count());
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);

  const int sel = (fold ? 4 : 0) | (gelu ? 2 : 0) | (res ? 1 : 0);
  switch (sel) {
    case 0: return dispatch_store<false, false, false>(tma_store, ta, tb, tc, p, grid, s);
    case 1: return dispatch_store<false, false, true>(tma_store, ta, tb, tc, p, grid, s);
    case 2: return dispatch_store<false, true, false>(tma_store, ta, tb, tc, p, grid, s);
    case 3: return dispatch_store<false, true, true>(tma_store, ta, tb, tc, p, grid, s);
    case 4: return dispatch_store<true, false, false>(tma_store, ta, tb, tc, p, grid, s);
    case 5: return dispatch_store<true, false, true>(tma_store, ta, tb, tc, p, grid, s);
    case 6: return dispatch_store<true, true, false>(tma_store, ta, tb, tc, p, grid, s);
    default: return dispatch_store<true, true, true>(tma_store, ta, tb, tc, p, grid, s);
  }
}
