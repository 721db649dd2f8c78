// Library-level C-ABI entry points: version and thread-local error string.
#include <stdarg.h>
#include <string.h>

#include "sad_common.cuh"

static thread_local char g_err[512] = "";

void sad_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int sad_version(void) { return SAD_ABI_VERSION; }

extern "C" const char* sad_last_error_string(void) { return g_err; }
