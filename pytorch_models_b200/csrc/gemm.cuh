// Persistent warp-specialised bf16 GEMM for sm_100a:  out[b][m][n] = epilogue( sum_k A[b][m][k] * W[n][k] )
//
//   warp 0      : TMA producer (A/W tiles -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (128 x 256 x 16 per instruction)
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> bias / LayerNorm-fold / erf-GELU / residual -> bf16 ->
//                 swizzled smem -> TMA store), overlapped with the next tile's main loop through two
//                 256-column TMEM accumulators.
//
// This one kernel serves every linear on the encoder path (reference call sites: transformer.py:47-49 fused QKV,
// transformer.py:53 out_proj, transformer.py:59-67 MLP, vit.py:78 patch embedding as a GEMM over patch rows).
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;  // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_STG_BYTES = 32 * 128;  // one staging buffer: 32 rows x 64 bf16
constexpr int GEMM_SMEM_RING = GEMM_STAGES * GEMM_STAGE_BYTES;
constexpr int GEMM_SMEM_STG = 4 * 2 * GEMM_STG_BYTES;     // 4 epilogue warps x 2 buffers
constexpr int GEMM_SMEM_COLVEC = 2 * GEMM_BN * 4;         // bias|c and colsum for one tile
constexpr int GEMM_SMEM_BYTES = GEMM_SMEM_RING + GEMM_SMEM_STG + GEMM_SMEM_COLVEC + 256;

struct GemmParams {
  int M;  // rows per batch
  int N;
  int K;
  int batches;
  int tiles_m;  // per batch
  int tiles_n;
  const float* bias;         // [N]  bias, or the folded constant vector c when ln_fold
  const float* colsum;       // [N]  s_n = sum_k W'[n][k] (ln_fold only)
  const float2* rowstats;    // [batches*M] (mean, rstd) (ln_fold only)
  const __nv_bfloat16* res;  // residual / positional table, nullptr if unused
  long long res_batch_stride;  // elements; 0 => same table for every batch (positional embedding)
  int ldr;                     // residual row stride (elements)
  __nv_bfloat16* out;          // used by the direct-store path only
  long long out_batch_stride;
  int ldo;
};

// erf-GELU: x*Phi(x) = relu(x) - |x| * 2^Q(|x|), Q = degree-6 minimax fit of log2(0.5*erfc(t/sqrt2)) on [0,6]
// (max abs error 2.8e-7 vs the exact erf form, measured in fp32; reference op: nn.GELU(), transformer.py:61).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float t = fminf(fabsf(x), 6.0f);
  float q = 3.310347528895363e-05f;
  q = fmaf(q, t, -0.0007693132502026856f);
  q = fmaf(q, t, 0.008081023581326008f);
  q = fmaf(q, t, -0.053412578999996185f);
  q = fmaf(q, t, -0.45877063274383545f);
  q = fmaf(q, t, -1.151201844215393f);
  q = fmaf(q, t, -0.9999930262565613f);
  return fmaf(-fabsf(x), fast_exp2(q), fmaxf(x, 0.0f));
}

template <bool kFold, bool kGelu, bool kRes, bool kTmaStore>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t ring = smem_base;
  const uint32_t stg_base = smem_base + GEMM_SMEM_RING;
  float* colvec = reinterpret_cast<float*>(smem + GEMM_SMEM_RING + GEMM_SMEM_STG);
  const uint32_t bars = smem_base + GEMM_SMEM_RING + GEMM_SMEM_STG + GEMM_SMEM_COLVEC;
  // barrier slots (8 bytes each): full[4] empty[4] tfull[2] tempty[2]; then the TMEM base address word.
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (GEMM_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * GEMM_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * GEMM_STAGES + 2 + a); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(
      smem + GEMM_SMEM_RING + GEMM_SMEM_STG + GEMM_SMEM_COLVEC + 8 * (2 * GEMM_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_base & 1023u) {
      printf("gemm_bf16_kernel: dynamic smem base not 1024-aligned (%u)\n", smem_base);
      __trap();
    }
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (kTmaStore) tma_prefetch_desc(&tmC);
  }
  if (warp == 1) {
    tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int tiles_per_batch = p.tiles_m * p.tiles_n;
  const int total_tiles = tiles_per_batch * p.batches;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch;
        const int r = tile - b * tiles_per_batch;
        const int m0 = (r / p.tiles_n) * GEMM_BM;
        const int n0 = (r % p.tiles_n) * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), GEMM_STAGE_BYTES);
          const uint32_t sa = ring + stage * GEMM_STAGE_BYTES;
          tma_load_3d(&tmA, full_bar(stage), sa, kb * GEMM_BK, m0, b);
          tma_load_2d(&tmB, full_bar(stage), sa + GEMM_A_BYTES, kb * GEMM_BK, n0);
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, GEMM_BN, 0, 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = ring + stage * GEMM_STAGE_BYTES;
          const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sa + GEMM_A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the 128B swizzle atom: +2 in the (addr >> 4) field
            umma_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == GEMM_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int e = (warp - 2) * 32 + lane;  // 0..127 among epilogue threads
    const uint32_t stg = stg_base + (warp - 2) * (2 * GEMM_STG_BYTES);
    uint8_t* stg_ptr = smem + GEMM_SMEM_RING + (warp - 2) * (2 * GEMM_STG_BYTES);
    float* cv_b = colvec;
    float* cv_s = colvec + GEMM_BN;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t buf = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_batch;
      const int r = tile - b * tiles_per_batch;
      const int m0 = (r / p.tiles_n) * GEMM_BM;
      const int n0 = (r % p.tiles_n) * GEMM_BN;

      named_bar_sync(1, 128);  // everyone finished reading the previous tile's column vectors
      for (int i = e; i < GEMM_BN; i += 128) {
        const int n = n0 + i;
        cv_b[i] = (n < p.N && p.bias != nullptr) ? __ldg(p.bias + n) : 0.0f;
        if (kFold) cv_s[i] = (n < p.N) ? __ldg(p.colsum + n) : 0.0f;
      }
      named_bar_sync(1, 128);

      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      float rstd = 1.0f, nmr = 0.0f;  // nmr = -mean * rstd
      if (kFold && row_ok) {
        const float2 st = __ldg(p.rowstats + (long long)b * p.M + row);
        rstd = st.y;
        nmr = -st.x * st.y;
      }
      const __nv_bfloat16* res_row = nullptr;
      if (kRes) res_row = p.res + (long long)b * p.res_batch_stride + (long long)row * p.ldr + n0;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * GEMM_BN + (uint32_t(q * 32) << 16);

#pragma unroll 1
      for (int c = 0; c < GEMM_BN / 64; ++c) {
        const int nc = n0 + c * 64;
        if (nc >= p.N) break;
        uint4 rres[8];
        if (kRes) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            rres[j] = make_uint4(0, 0, 0, 0);
            if (row_ok && nc + j * 8 < p.N) rres[j] = __ldg(reinterpret_cast<const uint4*>(res_row + c * 64) + j);
          }
        }
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + c * 64, v0);
        tmem_ld32(taddr + c * 64 + 32, v1);
        tmem_wait_ld();

        uint32_t packed[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x0, x1;
          {
            const int col = c * 64 + 2 * j;
            const float a0 = __uint_as_float(j < 16 ? v0[2 * j] : v1[2 * j - 32]);
            const float a1 = __uint_as_float(j < 16 ? v0[2 * j + 1] : v1[2 * j + 1 - 32]);
            if (kFold) {
              x0 = fmaf(rstd, a0, fmaf(nmr, cv_s[col], cv_b[col]));
              x1 = fmaf(rstd, a1, fmaf(nmr, cv_s[col + 1], cv_b[col + 1]));
            } else {
              x0 = a0 + cv_b[col];
              x1 = a1 + cv_b[col + 1];
            }
          }
          if (kGelu) {
            x0 = gelu_erf_fast(x0);
            x1 = gelu_erf_fast(x1);
          }
          if (kRes) {
            const uint32_t rr = reinterpret_cast<const uint32_t*>(rres)[j];
            x0 += bf16_lo(rr);
            x1 += bf16_hi(rr);
          }
          packed[j] = pack_bf16x2(x0, x1);
        }

        if (kTmaStore) {
          if (lane == 0) tma_store_wait_read<1>();  // the store that last used this buffer has drained
          __syncwarp();
          uint8_t* dst = stg_ptr + buf * GEMM_STG_BYTES + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            *reinterpret_cast<uint4*>(dst + ((j ^ (lane & 7)) << 4)) =
                make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tmC, stg + buf * GEMM_STG_BYTES, nc, m0 + q * 32, b);
            tma_store_commit();
          }
          buf ^= 1u;
        } else {
          if (row_ok) {
            __nv_bfloat16* orow = p.out + (long long)b * p.out_batch_stride + (long long)row * p.ldo + nc;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (nc + j * 8 < p.N)
                *(reinterpret_cast<uint4*>(orow) + j) =
                    make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            }
          }
        }
      }
      // all TMEM reads of this accumulator are complete (wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (kTmaStore && lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace b200
