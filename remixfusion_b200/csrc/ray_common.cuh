// ray_common.cuh — pieces shared by the Stage-2 kernels (ray_query.cu: fp32 SIMT decoder; ray_query_tc.cu: tcgen05
// tensor-core decoder): kernel-side configuration, sample position, hash / GBV feature gathers, tsdf clamps.
#pragma once
#include <math.h>
#include "grid_encode.cuh"

namespace rf {

constexpr int kGeo = 15;            // decoder.geo_feat_dim
constexpr int kNB = 16;             // pos.n_bins
constexpr int kBlob = 3 * kNB;      // 48
constexpr int kOut1 = 1 + kGeo;     // 16
constexpr int kIn2 = kBlob + kGeo + 3;   // 66
constexpr int kMaxS = 128;
constexpr int kTile = 128;          // samples per block iteration == threads per block

struct RayK {
    int   n_range_d, n_samples_d, S, perturb;
    float c_trunc, trunc, clamp_thr, sc_trunc, depth_trunc;
    int   clamp_mode, rgb_all_ones, n_hash_out, in1;      // in1 = n_hash_out + 48 + 1
    double b0[3], bl[3];                                   // bbox low corner and extent (float64, model/scene_rep.py:388)
    long long n_rays, n_total;                             // local rays, and rays in the whole (multi-GPU) batch
};

// ------------------------------------------------------------------------------------------------------------
// Per-sample building blocks
// ------------------------------------------------------------------------------------------------------------
template <int HID>
__device__ __forceinline__ void axpy_row(float (&h)[HID], float a, const float* __restrict__ w) {
#pragma unroll
    for (int j = 0; j < HID; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(w + j);
        h[j] = fmaf(a, v.x, h[j]); h[j + 1] = fmaf(a, v.y, h[j + 1]); h[j + 2] = fmaf(a, v.z, h[j + 2]); h[j + 3] = fmaf(a, v.w, h[j + 3]);
    }
}
template <int HID>
__device__ __forceinline__ float dot_row(const float (&h)[HID], const float* __restrict__ w) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < HID; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(w + j);
        a0 = fmaf(h[j], v.x, a0); a1 = fmaf(h[j + 1], v.y, a1); a2 = fmaf(h[j + 2], v.z, a2); a3 = fmaf(h[j + 3], v.w, a3);
    }
    return (a0 + a1) + (a2 + a3);
}

// sample position in normalised coordinates: pts = o + d*z (separate fp32 mul/add, :443), then float64
// normalisation (:388) and the cast to fp32 tiny-cuda-nn applies at its boundary.
__device__ __forceinline__ void sample_x(const RayK& k, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                         long long r, float z, float (&x)[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float p = __fadd_rn(__ldg(rays_o + 3 * r + a), __fmul_rn(__ldg(rays_d + 3 * r + a), z));
        x[a] = (float)(((double)p - k.b0[a]) / k.bl[a]);
    }
}

__device__ __forceinline__ float2 hash_level_feat(const GridDev& g, const float* __restrict__ params, int l, const float (&x)[3]) {
    unsigned cx, cy, cz; float fx, fy, fz;
    pos_fract(x[0], g.scale[l], cx, fx); pos_fract(x[1], g.scale[l], cy, fy); pos_fract(x[2], g.scale[l], cz, fz);
    const float2* tab = reinterpret_cast<const float2*>(params) + g.offset[l];
    const unsigned size = g.size[l], res = g.res[l];
    float2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        v[c] = __ldg(tab + grid_index(g.is_hash, size, res, cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)));
    float f0 = 0.f, f1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = corner_weight(c, fx, fy, fz);
        f0 = fmaf(w, v[c].x, f0); f1 = fmaf(w, v[c].y, f1);
    }
    return make_float2(f0, f1);
}

__device__ __forceinline__ float4 gbv_feat(const GridDev& g, const float* __restrict__ params, const float (&x)[3]) {
    unsigned cx, cy, cz; float fx, fy, fz;
    pos_fract(x[0], g.scale[0], cx, fx); pos_fract(x[1], g.scale[0], cy, fy); pos_fract(x[2], g.scale[0], cz, fz);
    const float4* tab = reinterpret_cast<const float4*>(params);
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        v[c] = __ldg(tab + grid_index(false, g.size[0], g.res[0], cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)));
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float w = corner_weight(c, fx, fy, fz);
        o.x = fmaf(w, v[c].x, o.x); o.y = fmaf(w, v[c].y, o.y); o.z = fmaf(w, v[c].z, o.z); o.w = fmaf(w, v[c].w, o.w);
    }
    return o;
}

// tsdf rescale and clamps — model/scene_rep.py:330-337 (and :230-233, :292-294 for the point-query variants)
// variant 0: as cfg.clamp_mode; 1: query_sdf_res (always +-1); 2: query_color_residual (decoder fed raw g0, nothing added)
__device__ __forceinline__ void tsdf_terms(const RayK& k, int variant, float g0, float& t_add, float& cin, float& dt_dg_add, float& dt_dg_cin) {
    float t = __fdiv_rn(__fmul_rn(g0, k.c_trunc), k.trunc);
    float s = k.c_trunc / k.trunc;
    if (variant == 2) { t_add = 0.f; cin = g0; dt_dg_add = 0.f; dt_dg_cin = 1.f; return; }
    if (variant == 0 && k.clamp_mode) {
        float thr = k.clamp_thr;
        t_add = fminf(fmaxf(t, -thr), thr);
        cin = fminf(fmaxf(t_add, -1.f), 1.f);
        dt_dg_add = (t >= -thr && t <= thr) ? s : 0.f;
        dt_dg_cin = (t_add >= -1.f && t_add <= 1.f) ? dt_dg_add : 0.f;
    } else {
        t_add = fminf(fmaxf(t, -1.f), 1.f);
        cin = t_add;
        dt_dg_add = (t >= -1.f && t <= 1.f) ? s : 0.f;
        dt_dg_cin = dt_dg_add;
    }
}

struct Weights { const float* w_sdf0; const float* w_sdf1; const float* w_col0; const float* w_col1; };

struct Grads { float* g_hash; float* g_w_sdf0; float* g_w_sdf1; float* g_w_col0; float* g_w_col1; };

// ---- Tensor-core path (mlp_precision 1): ray_encode.cu (operand tiles, table-gradient scatter) + ray_mlp_tc.cu ------------
// Workspace written by the forward and re-read by the backward (floats; P = n_rays * S samples, plane index q = s * n_rays + r,
// tile = 128 consecutive plane indices):
//   [0, 4096 NT)            hash features as READY tcgen05 operands, one 16 KB block per tile: bf16 hi parts of the four
//                           8-column chunks (2 KB each: row m at byte m * 16 = levels c, c+4, c+8, c+12 x 2 features), then the lo parts
//   [gbv, gbv + 4P)         GBV trilinear features [P] float4
//   [xn, xn + 3P)           normalised positions [3][P];   then (rays only) depth along the ray [P]
// Feature gradients (backward scratch): [4 chunks][P][8 floats], same column order.
// X-order of the hash features: operand chunk c (8 columns) holds levels c, c + 4, c + 8, c + 12 — one coarse, two middle and
// one fine level per chunk, so that the walking roles of ray_encode.cu (one chunk each) carry equal work.  Column
// 8 c + 2 j + f of X (and of W0, dX) <-> level c + 4 j, feature f, i.e. column 2 (c + 4 j) + f of the reference's layout.
__host__ __device__ inline int hash_col_to_feature(int kx) { return 2 * ((kx >> 3) + 4 * ((kx >> 1) & 3)) + (kx & 1); }
__host__ __device__ inline long long ws_tiles(long long P) { return (P + 127) >> 7; }
__host__ __device__ inline long long ws_off_gbv(long long P) { return ws_tiles(P) * 4096; }
__host__ __device__ inline long long ws_off_xn(long long P) { return ws_off_gbv(P) + 4 * P; }
inline long long ws_floats(long long P, bool with_z) { return ws_off_xn(P) + (with_z ? 4 : 3) * P; }
constexpr int kHopTileBytes = 16384;

bool tc_supported(const RayK& k, int hidden);
size_t scatter_scratch_floats(const GridDev& hg, long long n_rays);
int launch_fwd_tc(const RayK& k, int hidden, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o,
                  const float* rays_d, const float* z_vals, long long P, float* raw, float* feat, cudaStream_t s);
int launch_points_tc(RayK k, int hidden, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* x, long long n,
                     int variant, float* raw, float* feat, cudaStream_t s);
// n_live [n_rays]: samples s >= n_live[r] of ray r have an all-zero upstream gradient (composite_bwd_kernel)
int launch_bwd_tc(const RayK& k, int hidden, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, long long P, const float* feat,
                  const float* d_raw_tot, const int* n_live, float* dfeat, const Grads& gr, float* g_rays_o, float* g_rays_d, cudaStream_t s);
int launch_scatter_raygrad(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, long long P, const float* feat,
                           const float* dfeat, const float* dgb, const float* dxb, const int* n_live, float* g_hash, float* g_rep, float* g_o,
                           float* g_d, cudaStream_t s);

}  // namespace rf
