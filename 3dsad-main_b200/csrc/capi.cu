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

static unsigned long long g_launches = 0;   // kernels launched by this library (all threads)
void sad_count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
extern "C" unsigned long long sad_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int sad_version(void) { return SAD_ABI_VERSION; }

extern "C" const char* sad_last_error_string(void) { return g_err; }
