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

// ---- optional per-kernel timing (bench.py's live roofline measurement) -------------------------------------
// When enabled, the launchers bracket each hot kernel with a pair of CUDA events on the launching stream; the
// events are only read back by rf_profile_read.  Disabled (the default) it costs one predictable branch.
static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[RF_PROF_SLOTS][2];
static bool g_prof_have[RF_PROF_SLOTS];
static bool g_prof_init = false;

bool prof_enabled() { return g_prof_on; }
void prof_mark(int slot, int which, cudaStream_t s) {
    if (!g_prof_on || slot < 0 || slot >= RF_PROF_SLOTS) return;
    cudaEventRecord(g_prof_ev[slot][which], s);
    if (which == 1) g_prof_have[slot] = true;
}

}  // namespace rf

extern "C" int rf_profile_enable(int on) {
    using namespace rf;
    if (on && !g_prof_init) {
        for (int i = 0; i < RF_PROF_SLOTS; ++i) { cudaEventCreate(&g_prof_ev[i][0]); cudaEventCreate(&g_prof_ev[i][1]); }
        g_prof_init = true;
    }
    for (int i = 0; i < RF_PROF_SLOTS; ++i) g_prof_have[i] = false;
    g_prof_on = on != 0;
    return 0;
}

extern "C" int rf_profile_read(float* ms) {
    using namespace rf;
    RF_REQUIRE(ms, RF_E_NULL, "rf_profile_read: NULL");
    for (int i = 0; i < RF_PROF_SLOTS; ++i) {
        ms[i] = -1.f;
        if (g_prof_init && g_prof_have[i]) {
            cudaEventSynchronize(g_prof_ev[i][1]);
            float t = 0.f;
            if (cudaEventElapsedTime(&t, g_prof_ev[i][0], g_prof_ev[i][1]) == cudaSuccess) ms[i] = t;
        }
    }
    return 0;
}

extern "C" int rf_version(void) { return RF_ABI_VERSION; }
extern "C" const char* rf_last_error(void) { return rf::last_error_buf(); }
