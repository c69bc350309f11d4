// error.cu -- thread-local last-error string + ABI version.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void rbm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rbm_last_error(void) { return g_err; }
extern "C" int rbm_abi_version(void) { return RBM_ABI_VERSION; }
