/*
 * b200enc — C-ABI of the B200-native encoder-block hot path (libb200enc.so).
 *
 * The reference (gau-nernst/pytorch-models) has no FFI: its boundary for this path is the nn.Module protocol of
 * pytorch_models/transformer.py. Each entry point below replaces the ATen ops that one reference call site
 * dispatches (cited per function). All tensors are device pointers owned by the caller (PyTorch); activations and
 * weights are bf16 row-major, vectors and statistics fp32. Every call enqueues on `stream` (a cudaStream_t) and
 * returns without synchronising; nothing is allocated or freed by the library.
 *
 * Return value: 0 = ok; < 0 = argument / shape / alignment error; > 0 = cudaError_t or CUresult.
 * b200enc_last_error() returns a thread-local description of the last non-zero return.
 */
#ifndef B200ENC_H_
#define B200ENC_H_

#ifdef __cplusplus
extern "C" {
#endif

#define B200ENC_VERSION 100

int b200enc_version(void);
const char* b200enc_last_error(void);

/* flags for b200enc_linear */
#define B200ENC_LINEAR_GELU 1          /* exact (erf) GELU after bias: nn.GELU(), transformer.py:61 */
#define B200ENC_LINEAR_DIRECT_STORE 256 /* debug: st.global epilogue instead of the TMA-store epilogue */

/*
 * out[b][m][n] = epi( sum_k x[b][m][k] * w[n][k] )   for b < batches, m < M, n < N      (tcgen05 GEMM)
 *
 * Replaces nn.Linear at transformer.py:47-49 (q/k/v projections, fused as one [3d,d] weight), :53 (out_proj),
 * :59 (linear1 + nn.GELU :61), :66 (linear2), the residual adds at :125-126, and — with x = patch rows —
 * the patch-embedding Conv2d at image/vit.py:78 plus the positional add at :79.
 *
 *   epi(acc) = acc + bias[n]                                        (colsum == NULL)
 *            = rstd[m]*(acc - mean[m]*colsum[n]) + bias[n]           (LayerNorm folded into the GEMM: w must be
 *              gamma-scaled, colsum[n] = sum_k w[n][k], bias[n] = W.beta + b; rowstats = (mean, rstd) pairs)
 *   then GELU if flags & B200ENC_LINEAR_GELU, then + residual[b][m][n] if residual != NULL
 *   (res_batch_stride == 0 broadcasts one [M, N] table over the batch: the positional embedding).
 *
 * Strides are in elements. K and N must be multiples of 8; rows 16-byte aligned.
 */
int b200enc_linear(const void* x, long long x_batch_stride, int ldx, const void* w, int ldw, const float* bias,
                   const float* colsum, const float* rowstats, const void* residual, long long res_batch_stride,
                   int ldr, void* out, long long out_batch_stride, int ldo, int batches, int M, int N, int K,
                   int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200ENC_H_ */
