// ABI version + thread-local error string of libsummerclip_b200.
#include <stdarg.h>

#include "sc_common.cuh"

namespace sc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace sc

extern "C" {
int sc_version(void) { return SC_ABI_VERSION; }
const char* sc_last_error(void) { return sc::g_err; }
}
