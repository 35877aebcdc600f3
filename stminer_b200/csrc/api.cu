// Library identification and per-thread error string of the stminer_b200 C ABI.
#include <stdarg.h>

#include "stm_common.cuh"

namespace stm {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

}  // namespace stm

extern "C" int stm_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* stm_last_error_string(void) { return stm::g_last_error; }
