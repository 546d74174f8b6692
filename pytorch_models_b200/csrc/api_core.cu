// Version / error-reporting entry points of the C-ABI (include/b200enc.h).
#include "../../include/b200enc.h"
#include "host_util.h"

extern "C" int b200enc_version(void) { return B200ENC_VERSION; }
extern "C" const char* b200enc_last_error(void) { return b200::last_error_buf(); }
