#include <cstdlib>
#include "host_util.h"

#include <string.h>

#include <mutex>
#include <unordered_map>

namespace b200 {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// ---------------------------------------------------------------- tensor-map cache
namespace {
struct TmapKey {
  uint64_t w[14];  // base, rank|swizzle, dims[5], strides[4], box packed in 3 words
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t v : k.w) {
      h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
      h *= 0xff51afd7ed558ccdull;
    }
    return size_t(h ^ (h >> 32));
  }
};
constexpr size_t kTmapCacheMax = 8192;  // ~1.3 MB; a ViT-L forward touches ~400 distinct maps
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
unsigned long long g_tmap_hits = 0, g_tmap_misses = 0;
}  // namespace

TmapCacheStats tmap_cache_stats() {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  return TmapCacheStats{g_tmap_hits, g_tmap_misses};
}

int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes, int elem_bytes) {
  if (rank < 1 || rank > 5) return set_error(-1, "tensor map rank %d out of range", rank);
  if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0) return set_error(-1, "tensor base %p not 16-byte aligned", base);
  for (int i = 0; i + 1 < rank; ++i)
    if (strides_bytes[i] % 16 != 0)
      return set_error(-1, "tensor map stride %d = %llu bytes is not a multiple of 16", i,
                       (unsigned long long)strides_bytes[i]);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = reinterpret_cast<uintptr_t>(base);
  key.w[1] = uint64_t(rank) | (uint64_t(swizzle_bytes) << 8) | (uint64_t(elem_bytes) << 16);
  for (int i = 0; i < rank; ++i) key.w[2 + i] = dims[i];
  for (int i = 0; i + 1 < rank; ++i) key.w[7 + i] = strides_bytes[i];
  for (int i = 0; i < rank; ++i) key.w[11 + i / 2] |= uint64_t(box[i]) << (32 * (i & 1));
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      ++g_tmap_hits;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(-2, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, cuuint32_t(rank), const_cast<void*>(base), d, st, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(int(r),
                     "cuTensorMapEncodeTiled failed (%d): rank=%d dims=%llu,%llu,%llu,%llu,%llu box=%u,%u,%u,%u,%u "
                     "stride0=%llu swizzle=%d",
                     int(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                     rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0,
                     (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), swizzle_bytes);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
    ++g_tmap_misses;
  }
  return 0;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                   uint64_t row_stride, uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows,
                   int swizzle_bytes) {
  if ((row_stride * 2) % 16 != 0)
    return set_error(-1, "row stride %llu elements is not a multiple of 8", (unsigned long long)row_stride);
  const int rank = batches == 0 ? 2 : 3;
  if (rank == 3 && (batch_stride * 2) % 16 != 0)
    return set_error(-1, "batch stride %llu elements is not a multiple of 8", (unsigned long long)batch_stride);
  const uint64_t dims[3] = {inner, rows, batches == 0 ? 1 : batches};
  const uint64_t strides[2] = {row_stride * 2, batch_stride * 2};
  const uint32_t box[3] = {box_inner, box_rows, 1};
  return make_tmap_bf16_nd(out, base, rank, dims, strides, box, swizzle_bytes);
}

int make_tmap_u8(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batches, uint64_t row_stride,
                 uint64_t batch_stride, uint32_t box_inner, uint32_t box_rows, int swizzle_bytes) {
  if (row_stride % 16 != 0) return set_error(-1, "row stride %llu bytes is not a multiple of 16", (unsigned long long)row_stride);
  const int rank = batches == 0 ? 2 : 3;
  if (rank == 3 && batch_stride % 16 != 0)
    return set_error(-1, "batch stride %llu bytes is not a multiple of 16", (unsigned long long)batch_stride);
  const uint64_t dims[3] = {inner, rows, batches == 0 ? 1 : batches};
  const uint64_t strides[2] = {row_stride, batch_stride};
  const uint32_t box[3] = {box_inner, box_rows, 1};
  return make_tmap_bf16_nd(out, base, rank, dims, strides, box, swizzle_bytes, 1);
}

// ---------------------------------------------------------------- per-device facts
namespace {
constexpr int kMaxDevices = 64;
}

int sm_count() {
  static int counts[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = counts[dev];
  if (n) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  counts[dev] = n;
  return n;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("B200ENC_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int ensure_dynamic_smem(const void* kernel, int bytes) {
  struct Key {
    const void* k;
    int dev;
    bool operator==(const Key& o) const { return k == o.k && dev == o.dev; }
  };
  struct KeyHash {
    size_t operator()(const Key& x) const { return std::hash<const void*>()(x.k) ^ (size_t(x.dev) * 0x9E3779B1u); }
  };
  static std::mutex mu;
  static std::unordered_map<Key, int, KeyHash> done;
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(Key{kernel, dev});
    if (it != done.end() && it->second >= bytes) return 0;
  }
  B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  std::lock_guard<std::mutex> lk(mu);
  done[Key{kernel, dev}] = bytes;
  return 0;
}

// ---------------------------------------------------------------- abort word
static unsigned int* g_abort_host = nullptr;

unsigned int* abort_word() {
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    if (cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess && p != nullptr) {
      memset(p, 0, 64);
      g_abort_host = static_cast<unsigned int*>(p);
    } else {
      (void)cudaGetLastError();
    }
  });
  return g_abort_host;  // unified addressing: the host pointer of mapped memory is valid on every device
}

unsigned int read_abort_word(bool clear) {
  unsigned int* w = abort_word();
  if (!w) return 0;
  const unsigned int v = *reinterpret_cast<volatile unsigned int*>(w);
  if (clear) *reinterpret_cast<volatile unsigned int*>(w) = 0;
  return v;
}

int check_abort(const char* who) {
  const unsigned int v = read_abort_word(false);
  if (v == 0) return 0;
  return set_error(-3,
                   "%s: refused — an earlier libb200enc kernel gave up waiting on an on-chip barrier (code 0x%08x) and "
                   "its results are invalid; call b200enc_async_status(1) to acknowledge and clear",
                   who, v);
}

}  // namespace b200
