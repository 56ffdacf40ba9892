// rf_abi.cu — error plumbing and version of the C-ABI (include/rf_abi.h).
#include <stdarg.h>
#include "rf_common.cuh"

namespace rf {

static thread_local char g_err[512] = "";

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace rf

extern "C" int rf_version(void) { return RF_ABI_VERSION; }
extern "C" const char* rf_last_error(void) { return rf::last_error_buf(); }
