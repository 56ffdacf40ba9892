// encoders.cu — stand-alone encoder entry points: what `tcnn.Encoding.forward/backward` provided to the reference
// (model/encodings.py:33-51 HashGrid, :65-76 OneBlob; model/scene_rep.py:60-93 Dense GBV / GBW).  Used by the
// drop-in `get_encoder` modules, by `query_sdf_res(embed=True)` + `SLAM.smoothness` (mp_slam/slam.py:193-217) and by
// the point-query API.  The fused ray kernels (ray_query.cu) share the device code in grid_encode.cuh.
#include <math.h>
#include "grid_encode.cuh"

namespace rf {

// One thread per (sample, level): 8 corner gathers of F floats, F fma chains (Appendix B4).
template <int F>
__global__ void grid_fwd_kernel(GridDev g, const float* __restrict__ params, const float* __restrict__ x,
                                long long n, float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int l = blockIdx.y;
    if (i >= n) return;
    float px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
    unsigned cx, cy, cz; float fx, fy, fz;
    pos_fract(px, g.scale[l], cx, fx); pos_fract(py, g.scale[l], cy, fy); pos_fract(pz, g.scale[l], cz, fz);
    const float* tab = params + (size_t)g.offset[l] * F;
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        unsigned idx = grid_index(g.is_hash, g.size[l], g.res[l], cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1));
        float w = corner_weight(c, fx, fy, fz);
        const float* e = tab + (size_t)idx * F;
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] = fmaf(w, __ldg(e + f), acc[f]);
    }
    float* o = out + i * (long long)(g.n_levels * F) + l * F;
#pragma unroll
    for (int f = 0; f < F; ++f) o[f] = acc[f];
}

// Backward: scatter w*dout into the table gradient (B5); optional input gradient (B6).
template <int F>
__global__ void grid_bwd_kernel(GridDev g, const float* __restrict__ params, const float* __restrict__ x, long long n,
                                const float* __restrict__ dout, float* __restrict__ gparams, float* __restrict__ dx) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int l = blockIdx.y;
    if (i >= n) return;
    float px = x[3 * i], py = x[3 * i + 1], pz = x[3 * i + 2];
    unsigned cx, cy, cz; float fx, fy, fz;
    pos_fract(px, g.scale[l], cx, fx); pos_fract(py, g.scale[l], cy, fy); pos_fract(pz, g.scale[l], cz, fz);
    float d[F];
    const float* dop = dout + i * (long long)(g.n_levels * F) + l * F;
#pragma unroll
    for (int f = 0; f < F; ++f) d[f] = dop[f];
    float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        unsigned idx = grid_index(g.is_hash, g.size[l], g.res[l], cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1));
        size_t e = ((size_t)g.offset[l] + idx) * F;
        if (gparams) {
            float w = corner_weight(c, fx, fy, fz);
            if (F == 2) atomicAdd(reinterpret_cast<float2*>(gparams + e), make_float2(w * d[0], w * d[1]));
            else {
#pragma unroll
                for (int f = 0; f < F; ++f) atomicAdd(gparams + e + f, w * d[f]);
            }
        }
        if (dx) {
            float v = 0.f;
#pragma unroll
            for (int f = 0; f < F; ++f) v = fmaf(d[f], __ldg(params + e + f), v);
            float wx = (c & 1) ? fx : 1.f - fx, wy = (c & 2) ? fy : 1.f - fy, wz = (c & 4) ? fz : 1.f - fz;
            gx += ((c & 1) ? v : -v) * wy * wz;
            gy += ((c & 2) ? v : -v) * wx * wz;
            gz += ((c & 4) ? v : -v) * wx * wy;
        }
    }
    if (dx) {
        float s = g.scale[l];
        atomicAdd(dx + 3 * i, gx * s); atomicAdd(dx + 3 * i + 1, gy * s); atomicAdd(dx + 3 * i + 2, gz * s);
    }
}

template <int NB>
__global__ void oneblob_fwd_kernel(const float* __restrict__ x, long long n3, float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // one thread per (sample, coordinate)
    if (i >= n3) return;
    float o[NB];
    oneblob_coord<NB>(x[i], o);
#pragma unroll
    for (int k = 0; k < NB; ++k) out[i * NB + k] = o[k];
}

template <int NB>
__global__ void oneblob_bwd_kernel(const float* __restrict__ x, long long n3, const float* __restrict__ dout, float* __restrict__ dx) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n3) return;
    float g[NB];
    oneblob_coord_grad<NB>(x[i], g);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < NB; ++k) acc = fmaf(g[k], dout[i * NB + k], acc);
    dx[i] = acc;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_grid_desc_init(rf_grid_desc* d, int n_levels, int n_features, int is_hash, int log2_hashmap_size,
                                 int base_resolution, double per_level_scale) {
    RF_REQUIRE(d, RF_E_NULL, "rf_grid_desc_init: NULL descriptor");
    RF_REQUIRE(n_levels >= 1 && n_levels <= RF_MAX_LEVELS, RF_E_RANGE, "rf_grid_desc_init: n_levels %d not in [1,%d]", n_levels, RF_MAX_LEVELS);
    RF_REQUIRE(n_features == 1 || n_features == 2 || n_features == 4, RF_E_UNSUPPORTED, "rf_grid_desc_init: n_features %d", n_features);
    RF_REQUIRE(!is_hash || (log2_hashmap_size >= 1 && log2_hashmap_size <= 30), RF_E_RANGE, "rf_grid_desc_init: log2_hashmap_size %d", log2_hashmap_size);
    RF_REQUIRE(base_resolution >= 1, RF_E_RANGE, "rf_grid_desc_init: base_resolution %d", base_resolution);
    memset(d, 0, sizeof(*d));
    d->n_levels = n_levels; d->n_features = n_features; d->is_hash = is_hash ? 1 : 0;
    // Appendix B1 (tiny-cuda-nn GridEncodingTemplated constructor + grid_scale / grid_resolution)
    // log2 / exp2 are evaluated in double and rounded once to float, so that this table is reproducible on any host
    // (tiny-cuda-nn evaluates exp2f on the device, <= 2 ulp: its scale may differ from this one in the last bit).
    const float pls = (float)per_level_scale;            // JSON double narrowed to float
    const float log2s = (float)log2((double)pls);
    unsigned long long offset = 0;
    for (int l = 0; l < n_levels; ++l) {
        float scale = (float)exp2((double)((float)l * log2s)) * (float)base_resolution - 1.0f;
        unsigned res = (unsigned)ceilf(scale) + 1u;
        const unsigned max_params = 0xFFFFFFFFu / 2;
        unsigned long long n = (powf((float)res, 3.0f) > (float)max_params) ? max_params : (unsigned long long)res * res * res;
        n = (n + 7ull) / 8ull * 8ull;
        if (is_hash) { unsigned long long cap = 1ull << log2_hashmap_size; if (n > cap) n = cap; }
        d->scale[l] = scale; d->resolution[l] = res; d->size[l] = (unsigned)n; d->offset[l] = (unsigned)offset;
        offset += n;
        RF_REQUIRE(offset < (1ull << 32), RF_E_UNSUPPORTED, "rf_grid_desc_init: table exceeds 2^32 entries");
    }
    d->offset[n_levels] = (unsigned)offset;
    return 0;
}

static int check_grid(const rf_grid_desc* d, const char* who) {
    RF_REQUIRE(d, RF_E_NULL, "%s: NULL descriptor", who);
    RF_REQUIRE(d->n_levels >= 1 && d->n_levels <= RF_MAX_LEVELS, RF_E_RANGE, "%s: bad n_levels", who);
    RF_REQUIRE(d->n_features == 1 || d->n_features == 2 || d->n_features == 4, RF_E_UNSUPPORTED, "%s: bad n_features", who);
    return 0;
}

extern "C" int rf_grid_encode_forward(const rf_grid_desc* d, const float* params, const float* x, int64_t n, float* out, void* stream) {
    int rc = check_grid(d, "rf_grid_encode_forward"); if (rc) return rc;
    RF_REQUIRE(n >= 0, RF_E_RANGE, "rf_grid_encode_forward: negative n");
    if (n == 0) return 0;
    RF_REQUIRE(params && x && out, RF_E_NULL, "rf_grid_encode_forward: NULL pointer");
    GridDev g = to_dev(d);
    dim3 grid((unsigned)((n + 255) / 256), d->n_levels);
    cudaStream_t s = (cudaStream_t)stream;
    if (d->n_features == 1) grid_fwd_kernel<1><<<grid, 256, 0, s>>>(g, params, x, n, out);
    else if (d->n_features == 2) grid_fwd_kernel<2><<<grid, 256, 0, s>>>(g, params, x, n, out);
    else grid_fwd_kernel<4><<<grid, 256, 0, s>>>(g, params, x, n, out);
    RF_CHECK_LAUNCH("rf_grid_encode_forward");
    return 0;
}

extern "C" int rf_grid_encode_backward(const rf_grid_desc* d, const float* params, const float* x, int64_t n,
                                       const float* dout, float* grad_params, float* dx, void* stream) {
    int rc = check_grid(d, "rf_grid_encode_backward"); if (rc) return rc;
    RF_REQUIRE(n >= 0, RF_E_RANGE, "rf_grid_encode_backward: negative n");
    if (n == 0) return 0;
    RF_REQUIRE(params && x && dout, RF_E_NULL, "rf_grid_encode_backward: NULL pointer");
    RF_REQUIRE(grad_params || dx, RF_E_NULL, "rf_grid_encode_backward: nothing to compute");
    RF_REQUIRE(d->n_features != 2 || !grad_params || ((uintptr_t)grad_params & 7) == 0, RF_E_ALIGN, "rf_grid_encode_backward: grad_params must be 8-byte aligned");
    GridDev g = to_dev(d);
    cudaStream_t s = (cudaStream_t)stream;
    if (dx) { cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * 3 * (size_t)n, s); if (e != cudaSuccess) return set_error((int)e, "memset dx: %s", cudaGetErrorString(e)); }
    dim3 grid((unsigned)((n + 255) / 256), d->n_levels);
    if (d->n_features == 1) grid_bwd_kernel<1><<<grid, 256, 0, s>>>(g, params, x, n, dout, grad_params, dx);
    else if (d->n_features == 2) grid_bwd_kernel<2><<<grid, 256, 0, s>>>(g, params, x, n, dout, grad_params, dx);
    else grid_bwd_kernel<4><<<grid, 256, 0, s>>>(g, params, x, n, dout, grad_params, dx);
    RF_CHECK_LAUNCH("rf_grid_encode_backward");
    return 0;
}

extern "C" int rf_oneblob_forward(const float* x, int64_t n, int n_bins, float* out, void* stream) {
    RF_REQUIRE(n >= 0, RF_E_RANGE, "rf_oneblob_forward: negative n");
    RF_REQUIRE(n_bins == 16, RF_E_UNSUPPORTED, "rf_oneblob_forward: n_bins %d (only pos.n_bins = 16 is built)", n_bins);
    if (n == 0) return 0;
    RF_REQUIRE(x && out, RF_E_NULL, "rf_oneblob_forward: NULL pointer");
    long long n3 = 3 * (long long)n;
    oneblob_fwd_kernel<16><<<(unsigned)((n3 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, n3, out);
    RF_CHECK_LAUNCH("rf_oneblob_forward");
    return 0;
}

extern "C" int rf_oneblob_backward(const float* x, int64_t n, int n_bins, const float* dout, float* dx, void* stream) {
    RF_REQUIRE(n >= 0, RF_E_RANGE, "rf_oneblob_backward: negative n");
    RF_REQUIRE(n_bins == 16, RF_E_UNSUPPORTED, "rf_oneblob_backward: n_bins %d", n_bins);
    if (n == 0) return 0;
    RF_REQUIRE(x && dout && dx, RF_E_NULL, "rf_oneblob_backward: NULL pointer");
    long long n3 = 3 * (long long)n;
    oneblob_bwd_kernel<16><<<(unsigned)((n3 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x, n3, dout, dx);
    RF_CHECK_LAUNCH("rf_oneblob_backward");
    return 0;
}
