// microbench.cu — the two denominators of the gather-bound roofline: random 8-byte loads and random
// fp32 reductions (RED.ADD) over a table, all SMs.  Timed with CUDA events on the given stream.
#include "rf_common.cuh"

namespace rf {

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s; }
__device__ __forceinline__ unsigned mix(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void gather_kernel(const float2* __restrict__ table, unsigned n_entries, long long ops_per_thread, float* sink) {
    unsigned s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    float acc = 0.f;
    for (long long i = 0; i < ops_per_thread; i += 8) {
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldg(table + mix(lcg(s)) % n_entries);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += v[k].x + v[k].y;
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void atomic_kernel(float* __restrict__ table, unsigned n_entries, long long ops_per_thread) {
    unsigned s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
    for (long long i = 0; i < ops_per_thread; ++i) atomicAdd(table + mix(lcg(s)) % n_entries, 1.0f);
}

template <typename F>
static int timed(F launch, int iters, float* ms, cudaStream_t s) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();                                      // warm-up
    cudaEventRecord(a, s);
    for (int i = 0; i < iters; ++i) launch();
    cudaEventRecord(b, s);
    cudaError_t e = cudaEventSynchronize(b);
    if (e == cudaSuccess) e = cudaGetLastError();
    float t = 0.f;
    cudaEventElapsedTime(&t, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    if (e != cudaSuccess) return set_error((int)e, "microbench: %s", cudaGetErrorString(e));
    *ms = t / (float)iters;
    return 0;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_microbench_gather(void* table, int64_t table_bytes, int64_t n_ops, int iters, float* ms, void* stream) {
    RF_REQUIRE(table && ms, RF_E_NULL, "rf_microbench_gather: NULL");
    RF_REQUIRE(table_bytes >= 8 && n_ops > 0 && iters > 0 && table_bytes / 8 < (1ll << 32), RF_E_RANGE, "rf_microbench_gather: bad sizes");
    int blocks = num_sms() * 8, threads = 256;
    long long per = (n_ops + (long long)blocks * threads - 1) / ((long long)blocks * threads);
    per = (per + 7) / 8 * 8;
    cudaStream_t s = (cudaStream_t)stream;
    float* sink = (float*)table;
    return timed([&] { gather_kernel<<<blocks, threads, 0, s>>>((const float2*)table, (unsigned)(table_bytes / 8), per, sink); }, iters, ms, s);
}

extern "C" int rf_microbench_atomic(void* table, int64_t table_bytes, int64_t n_ops, int iters, float* ms, void* stream) {
    RF_REQUIRE(table && ms, RF_E_NULL, "rf_microbench_atomic: NULL");
    RF_REQUIRE(table_bytes >= 4 && n_ops > 0 && iters > 0 && table_bytes / 4 < (1ll << 32), RF_E_RANGE, "rf_microbench_atomic: bad sizes");
    int blocks = num_sms() * 8, threads = 256;
    long long per = (n_ops + (long long)blocks * threads - 1) / ((long long)blocks * threads);
    cudaStream_t s = (cudaStream_t)stream;
    return timed([&] { atomic_kernel<<<blocks, threads, 0, s>>>((float*)table, (unsigned)(table_bytes / 4), per); }, iters, ms, s);
}
