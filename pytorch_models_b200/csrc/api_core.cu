// Version / error-reporting entry points of the C-ABI (include/b200enc.h).
#include "../../include/b200enc.h"
#include "host_util.h"

extern "C" int b200enc_version(void) { return B200ENC_VERSION; }
extern "C" const char* b200enc_last_error(void) { return b200::last_error_buf(); }
extern "C" unsigned int b200enc_async_status(int clear) { return b200::read_abort_word(clear != 0); }
extern "C" void b200enc_tensor_map_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  const b200::TmapCacheStats s = b200::tmap_cache_stats();
  if (hits) *hits = s.hits;
  if (misses) *misses = s.misses;
}
