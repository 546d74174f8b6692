// Stand-alone GPU self-test / micro-benchmark for libb200enc (no PyTorch): each case checks a kernel against a
// straightforward fp64 CPU computation on the same inputs. Test infrastructure only; run on the B200 box:
//   ./b200enc_selftest <case> [args]      (exit code 0 = pass)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b200enc.h"

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e = (x);                                                              \
    if (e != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e)); \
      exit(3);                                                                        \
    }                                                                                 \
  } while (0)

static uint64_t g_seed = 0x9E3779B97F4A7C15ull;
static inline uint32_t rnd() {
  g_seed = g_seed * 6364136223846793005ull + 1442695040888963407ull;
  return uint32_t(g_seed >> 33);
}
static inline float urand() { return (rnd() & 0xFFFFFF) / float(0x1000000) * 2.0f - 1.0f; }  // [-1, 1)

static inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return uint16_t(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  explicit DevBuf(size_t b) : bytes(b) { CK(cudaMalloc(&p, b ? b : 16)); }
  ~DevBuf() { cudaFree(p); }
};

static std::vector<uint16_t> rand_bf16(size_t n, float scale, bool ints) {
  std::vector<uint16_t> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = f2bf(ints ? float(int(rnd() % 5) - 2) : urand() * scale);
  return v;
}
static std::vector<float> rand_f32(size_t n, float scale) {
  std::vector<float> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = urand() * scale;
  return v;
}

static double gelu_ref(double x) { return 0.5 * x * (1.0 + erf(x / sqrt(2.0))); }

struct CmpStat {
  double max_abs = 0, max_rel = 0;
  long long bad = 0, n = 0;
  int shown = 0;
  long long bad_rowmod8[8] = {0}, bad_colblk[8] = {0};
};
static void cmp_one(CmpStat& st, double want, float got, double atol, double rtol, int r, int c, const char* tag) {
  double d = fabs(double(got) - want);
  st.n++;
  if (!(d <= st.max_abs)) st.max_abs = (d != d) ? INFINITY : (d > st.max_abs ? d : st.max_abs);
  double rel = d / (fabs(want) + 1e-6);
  if (rel > st.max_rel && fabs(want) > 1e-2) st.max_rel = rel;
  if (!(d <= atol + rtol * fabs(want))) {
    st.bad++;
    st.bad_rowmod8[r & 7]++;
    st.bad_colblk[(c >> 3) & 7]++;
    if (st.shown < 12) {
      printf("    MISMATCH %s row=%d col=%d got=%.6g want=%.6g\n", tag, r, c, got, want);
      st.shown++;
    }
  }
}
static bool report(const char* name, const CmpStat& st) {
  printf("  [%s] %s: checked=%lld bad=%lld max_abs=%.4g max_rel=%.4g\n", st.bad ? "FAIL" : "ok", name, st.n, st.bad,
         st.max_abs, st.max_rel);
  if (st.bad) {
    printf("    bad by row%%8:");
    for (int i = 0; i < 8; ++i) printf(" %lld", st.bad_rowmod8[i]);
    printf("   by (col/8)%%8:");
    for (int i = 0; i < 8; ++i) printf(" %lld", st.bad_colblk[i]);
    printf("\n");
  }
  fflush(stdout);
  return st.bad == 0;
}

// ------------------------------------------------------------------------------------------ linear
struct LinearCase {
  const char* name;
  int batches, M, N, K;
  bool fold, gelu, res, res_bcast, direct, ints;
  int out_row_off;  // output rows shifted by this inside a (M + off)-row batch (patch-embed layout)
  int check_rows;   // rows per batch checked on the CPU (0 = all)
  int time_iters;
};

static bool run_linear(const LinearCase& c) {
  printf("linear %s: batches=%d M=%d N=%d K=%d fold=%d gelu=%d res=%d bcast=%d direct=%d\n", c.name, c.batches, c.M,
         c.N, c.K, c.fold, c.gelu, c.res, c.res_bcast, c.direct);
  fflush(stdout);
  const int B = c.batches, M = c.M, N = c.N, K = c.K;
  const int Mo = M + c.out_row_off;
  const float xs = c.ints ? 1.f : 1.0f, ws = c.ints ? 1.f : 0.05f;
  auto hx = rand_bf16(size_t(B) * M * K, xs, c.ints);
  auto hw = rand_bf16(size_t(N) * K, ws, c.ints);
  auto hb = rand_f32(N, c.ints ? 0.f : 0.5f);
  std::vector<float> hs(N, 0.f), hstats(size_t(B) * M * 2, 0.f);
  if (c.fold) {
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += bf2f(hw[size_t(n) * K + k]);
      hs[n] = float(s);
    }
    for (size_t i = 0; i < size_t(B) * M; ++i) {
      hstats[2 * i] = urand() * 0.3f;
      hstats[2 * i + 1] = 0.5f + fabsf(urand());
    }
  }
  const size_t res_rows = c.res_bcast ? size_t(M) : size_t(B) * M;
  auto hr = rand_bf16(c.res ? res_rows * N : 1, 1.0f, c.ints);

  DevBuf dx(hx.size() * 2), dw(hw.size() * 2), db(N * 4), ds(N * 4), dst(hstats.size() * 4), dr(hr.size() * 2),
      dout(size_t(B) * Mo * N * 2);
  CK(cudaMemcpy(dx.p, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw.p, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db.p, hb.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ds.p, hs.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dst.p, hstats.data(), hstats.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dr.p, hr.data(), hr.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout.p, 0x7f, dout.bytes));  // sentinel 0x7f7f = 3.39e38 in bf16

  const int flags = (c.gelu ? B200ENC_LINEAR_GELU : 0) | (c.direct ? B200ENC_LINEAR_DIRECT_STORE : 0);
  auto call = [&]() {
    return b200enc_linear(dx.p, (long long)M * K, K, dw.p, K, (const float*)db.p, c.fold ? (const float*)ds.p : nullptr,
                          c.fold ? (const float*)dst.p : nullptr, c.res ? dr.p : nullptr,
                          c.res_bcast ? 0 : (long long)M * N, N,
                          (uint16_t*)dout.p + size_t(c.out_row_off) * N, (long long)Mo * N, N, B, M, N, K, flags,
                          nullptr);
  };
  int rc = call();
  if (rc) {
    printf("  [FAIL] b200enc_linear rc=%d: %s\n", rc, b200enc_last_error());
    return false;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  [FAIL] kernel error: %s\n", cudaGetErrorString(e));
    return false;
  }
  std::vector<uint16_t> ho(size_t(B) * Mo * N);
  CK(cudaMemcpy(ho.data(), dout.p, ho.size() * 2, cudaMemcpyDeviceToHost));

  CmpStat st;
  const int step = (c.check_rows > 0 && c.check_rows < M) ? M / c.check_rows : 1;
  for (int b = 0; b < B; ++b) {
    for (int m = 0; m < M; m += 1) {
      // always check the first and last 130 rows (tile edges), sample the middle
      if (!(m < 130 || m >= M - 130 || (m % step) == 0)) continue;
      const uint16_t* xr = &hx[(size_t(b) * M + m) * K];
      for (int n = 0; n < N; ++n) {
        const uint16_t* wr = &hw[size_t(n) * K];
        double acc = 0;
        for (int k = 0; k < K; ++k) acc += double(bf2f(xr[k])) * double(bf2f(wr[k]));
        double v;
        if (c.fold) {
          const double mean = hstats[2 * (size_t(b) * M + m)], rstd = hstats[2 * (size_t(b) * M + m) + 1];
          v = rstd * (acc - mean * hs[n]) + hb[n];
        } else {
          v = acc + hb[n];
        }
        if (c.gelu) v = gelu_ref(v);
        if (c.res) v += bf2f(hr[(c.res_bcast ? size_t(m) : size_t(b) * M + m) * N + n]);
        const float got = bf2f(ho[(size_t(b) * Mo + c.out_row_off + m) * N + n]);
        cmp_one(st, v, got, c.ints ? 1e-6 : 0.02, c.ints ? 0.0 : 0.01, b * M + m, n, c.name);
      }
    }
    // rows before out_row_off must keep the sentinel
    for (int m = 0; m < c.out_row_off; ++m)
      for (int n = 0; n < N; ++n)
        if (ho[(size_t(b) * Mo + m) * N + n] != 0x7f7f) {
          st.bad++;
          if (st.shown++ < 12) printf("    CLOBBERED guard row b=%d m=%d n=%d\n", b, m, n);
        }
  }
  bool ok = report(c.name, st);

  if (ok && c.time_iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) call();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.time_iters; ++i) call();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.time_iters;
    const double fl = 2.0 * B * double(M) * N * K;
    printf("  time %s: %.3f ms  %.1f TFLOP/s\n", c.name, ms, fl / ms * 1e-9);
    fflush(stdout);
  }
  return ok;
}

static const LinearCase kLinearCases[] = {
    //  name            B   M      N     K    fold   gelu   res    bcast  direct ints  off rows iters
    {"min_direct", 1, 128, 256, 64, false, false, false, false, true, true, 0, 0, 0},
    {"min_tma", 1, 128, 256, 64, false, false, false, false, false, true, 0, 0, 0},
    {"k256_direct", 1, 128, 256, 256, false, false, false, false, true, true, 0, 0, 0},
    {"k256_tma", 1, 128, 256, 256, false, false, false, false, false, true, 0, 0, 0},
    {"multi_tile", 1, 512, 768, 768, false, false, false, false, false, false, 0, 0, 0},
    {"tails_direct", 1, 300, 576, 192, false, false, false, false, true, false, 0, 0, 0},
    {"tails_tma", 1, 300, 576, 192, false, false, false, false, false, false, 0, 0, 0},
    {"many_tiles", 1, 19000, 768, 768, false, false, false, false, false, false, 0, 40, 0},
    {"embed_like", 3, 196, 768, 768, false, false, true, true, false, false, 1, 0, 0},
    {"fold_gelu", 1, 1000, 3072, 768, true, true, false, false, false, false, 0, 60, 0},
    {"residual", 1, 1000, 768, 3072, false, false, true, false, false, false, 0, 60, 0},
    {"fold_qkv", 1, 1000, 2304, 768, true, false, false, false, false, false, 0, 60, 0},
    {"perf_qkv", 1, 25216, 2304, 768, true, false, false, false, false, false, 0, 16, 20},
    {"perf_out", 1, 25216, 768, 768, false, false, true, false, false, false, 0, 16, 20},
    {"perf_fc1", 1, 25216, 3072, 768, true, true, false, false, false, false, 0, 16, 20},
    {"perf_fc2", 1, 25216, 768, 3072, false, false, true, false, false, false, 0, 16, 20},
    {"perf_qkv_direct", 1, 25216, 2304, 768, true, false, false, false, true, false, 0, 16, 20},
};

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("usage: %s <case>|list\n", argv[0]);
    return 2;
  }
  std::string which = argv[1];
  if (which == "list") {
    for (auto& c : kLinearCases) printf("linear:%s\n", c.name);
    return 0;
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  bool ok = true, found = false;
  for (auto& c : kLinearCases) {
    if (which == std::string("linear:") + c.name || which == "linear:all") {
      found = true;
      ok = run_linear(c) && ok;
    }
  }
  if (!found) {
    printf("unknown case %s\n", which.c_str());
    return 2;
  }
  printf("%s\n", ok ? "SELFTEST PASS" : "SELFTEST FAIL");
  return ok ? 0 : 1;
}
