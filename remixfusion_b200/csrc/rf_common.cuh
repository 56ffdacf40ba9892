// rf_common.cuh — shared helpers for the librf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rf_abi.h"

namespace rf {

// Thread-local error string behind rf_last_error().
char* last_error_buf();
int   set_error(int code, const char* fmt, ...);

#define RF_REQUIRE(cond, code, ...)                      \
    do {                                                 \
        if (!(cond)) return rf::set_error((code), __VA_ARGS__); \
    } while (0)

#define RF_CHECK_LAUNCH(what)                                                         \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess)                                                       \
            return rf::set_error((int)e__, "%s: %s", (what), cudaGetErrorString(e__)); \
    } while (0)

// per-kernel event timing (rf_profile_enable / rf_profile_read)
bool prof_enabled();
void prof_mark(int slot, int which, cudaStream_t s);
struct ProfScope {
    int slot; cudaStream_t s;
    ProfScope(int slot_, cudaStream_t s_) : slot(slot_), s(s_) { prof_mark(slot, 0, s); }
    ~ProfScope() { prof_mark(slot, 1, s); }
};

static inline int num_sms() {                      // per device (a process may drive several GPUs); no cached static
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rf
