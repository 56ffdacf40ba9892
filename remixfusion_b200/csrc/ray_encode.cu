// ray_encode.cu — the gather / scatter half of the Stage-2 tensor-core path (mlp_precision 1).
//
//   ray_pos_kernel      pts = o + d*z, float64 normalisation by the bounding box (model/scene_rep.py:443,388) -> xn planes
//   encode_walk_kernel  hash levels + GBV trilinear features (tiny-cuda-nn grid forward, SURVEY Appendix B2-B4;
//                       model/scene_rep.py:325,329) -> feature planes
//   scatter_walk_kernel hash-table gradient (Appendix B5; autograd of model/scene_rep.py:325)
//
// Both grid kernels run one thread per (ray, level) that WALKS the samples of its ray in order.  Samples along a ray are
// sorted in depth (model/scene_rep.py:428), so consecutive samples stay in the same grid cell for a while on all but
// the finest levels: the forward keeps the 8 corner values in registers until the cell changes (one gather per cell
// run instead of one per sample); the backward accumulates the 8 corner gradients in registers and issues its 8
// vector reductions (RED.ADD.F32x2) only when the cell changes.  All planes are SAMPLE-MAJOR (index s * n_rays + r),
// so the lanes of a warp (consecutive rays) read and write consecutive addresses at every step of the walk.  The
// per-sample arithmetic (pos = fma(scale, x, 0.5), corner order, fma accumulation order) is the one grid_encode.cuh
// uses everywhere, so the features are bit-identical to the thread-per-sample fp32 kernels.
//
//   raygrad_walk_kernel BA mode (model/scene_rep.py:443 under mp_slam/mapper.py:456,484-485): dL/d rays_o, dL/d rays_d
//                       through the trilinear weights of the hash levels and the GBV (Appendix B6) + the OneBlob part
//
// Workspace layout (floats), P = n_rays * S:   [0, 2L*P) hash features [L][S][N][2];  [2L*P, 2L*P + 4P) GBV [S][N][4];
//                                              then xn [3][S][N]; then (rays only) z [S][N].
#include <stdlib.h>
#include <algorithm>
#include "ray_common.cuh"

namespace rf {

// Plane index of sample s of ray r: sample-major, so that threads that each walk one ray touch consecutive addresses.
//   q = s * n_rays + r
// Positions: pts = o + d*z (:443), float64 normalisation (:388), written through a shared-memory transpose so that both
// the [N][S] read of z_vals and the [S][N] write of the planes are full 128-byte segments.  Block = 32 rays.
__global__ void __launch_bounds__(256) ray_pos_kernel(RayK k, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                      const float* __restrict__ z_vals, long long P, float* __restrict__ xn) {
    extern __shared__ float sx[];                       // [4][S][33]: x, y, z of the normalised position, depth along the ray
    const int S = k.S;
    const long long r0 = blockIdx.x * 32ll;
    const int nr = (int)min(32ll, k.n_rays - r0);
    for (int i = threadIdx.x; i < nr * S; i += 256) {
        const int rl = i / S, s = i - rl * S;
        float x[3];
        const float zv = z_vals[r0 * S + i];
        sample_x(k, rays_o, rays_d, r0 + rl, zv, x);
        sx[s * 33 + rl] = x[0]; sx[(S + s) * 33 + rl] = x[1]; sx[(2 * S + s) * 33 + rl] = x[2]; sx[(3 * S + s) * 33 + rl] = zv;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 32 * S; j += 256) {
        const int s = j >> 5, rl = j & 31;
        if (rl < nr) {
            const long long q = (long long)s * k.n_rays + r0 + rl;
            xn[q] = sx[s * 33 + rl]; xn[P + q] = sx[(S + s) * 33 + rl]; xn[2 * P + q] = sx[(2 * S + s) * 33 + rl];
            xn[3 * P + q] = sx[(3 * S + s) * 33 + rl];
        }
    }
}

// Point queries (model/scene_rep.py:212-310): positions are given, already normalised: [n][3] -> planes [3][n].
__global__ void __launch_bounds__(256) point_pos_kernel(const float* __restrict__ x, long long n, float* __restrict__ xn) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    xn[i] = x[3 * i]; xn[n + i] = x[3 * i + 1]; xn[2 * n + i] = x[3 * i + 2];
}

// Features.  Thread = (ray, level), blockIdx.y + level0 = level (0..L-1 hash levels, L = GBV); lanes = consecutive rays.
__global__ void __launch_bounds__(128) encode_walk_kernel(GridDev hg, GridDev gg, const float* __restrict__ hash_params,
                                                          const float* __restrict__ gbv_params, const float* __restrict__ xn,
                                                          long long P, long long N, int S, int seg, int level0, float* __restrict__ feat) {
    const int l = blockIdx.y + level0, L = hg.n_levels;
    // unit = (ray, segment of `seg` samples): lanes are consecutive rays of one segment.  Large batches use one segment
    // (the whole ray); small ones are cut so that the dependent walk is short and the grid fills the machine.
    const long long unit = blockIdx.x * 128ll + threadIdx.x;
    const long long r = unit % N;
    const int s_begin = (int)(unit / N) * seg;
    if (s_begin >= S) return;
    const int s_end = min(S, s_begin + seg);
    const long long first = (long long)s_begin * N + r;
    const float* xs = xn + first; const float* ys = xn + P + first; const float* zs = xn + 2 * P + first;
    unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false;
    float xa = __ldg(xs), ya = __ldg(ys), za = __ldg(zs);
    CornerIndexer ci;
    unsigned idx[8];
    if (l < L) {
        const float scale = hg.scale[l];
        ci.init(hg.is_hash != 0, hg.size[l], hg.res[l]);
        const float2* tab = reinterpret_cast<const float2*>(hash_params) + hg.offset[l];
        float2* out = reinterpret_cast<float2*>(feat) + (long long)l * P + first;
        float2 v[8];
        for (int s = s_begin; s < s_end; ++s) {
            const float x = xa, y = ya, z = za;
            if (s + 1 < s_end) { xs += N; ys += N; zs += N; xa = __ldg(xs); ya = __ldg(ys); za = __ldg(zs); }
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
                ci.cell(cx, cy, cz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + idx[c]);
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float f0 = 0.f, f1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float w = corner_weight(c, fx, fy, fz);
                f0 = fmaf(w, v[c].x, f0); f1 = fmaf(w, v[c].y, f1);
            }
            *out = make_float2(f0, f1);
            out += N;
        }
    } else {
        const float scale = gg.scale[0];
        ci.init(false, gg.size[0], gg.res[0]);
        const float4* tab = reinterpret_cast<const float4*>(gbv_params);
        float4* out = reinterpret_cast<float4*>(feat + 2ll * L * P) + first;
        float4 v[8];
        for (int s = s_begin; s < s_end; ++s) {
            const float x = xa, y = ya, z = za;
            if (s + 1 < s_end) { xs += N; ys += N; zs += N; xa = __ldg(xs); ya = __ldg(ys); za = __ldg(zs); }
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
                ci.cell(cx, cy, cz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + idx[c]);
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float w = corner_weight(c, fx, fy, fz);
                o4.x = fmaf(w, v[c].x, o4.x); o4.y = fmaf(w, v[c].y, o4.y); o4.z = fmaf(w, v[c].z, o4.z); o4.w = fmaf(w, v[c].w, o4.w);
            }
            *out = o4;
            out += N;
        }
    }
}

// Table-gradient scatter, run-length reduced.  dfeat [L][P][2] (sample-major planes); thread = (ray, level): each lane
// accumulates the 8 corner gradients of its current cell in registers and issues the 8 vector reductions
// (RED.ADD.F32x2) when its cell changes; runs whose gradients are all zero (samples past the truncation mask) issue none.
// Small (coarse) levels are the contended ones: every ray of the batch lands on the same few thousand entries, and
// reductions onto one L2 line serialise.  Those levels accumulate into K private replicas of their gradient table
// (replica = block index mod K, so that neighbouring blocks — neighbouring pixels — never share one) which
// replica_reduce_kernel folds into the caller's table afterwards.
// d/dx of the trilinear interpolant of the corner scalars u[c] (c = cx + 2 cy + 4 cz): differences along one axis,
// bilinear weights of the other two (Appendix B6)
__device__ __forceinline__ void tri_grad(const float (&u)[8], float fx, float fy, float fz, float& gx, float& gy, float& gz) {
    const float ax = 1.f - fx, ay = 1.f - fy, az = 1.f - fz;
    const float w00 = ay * az, w10 = fy * az, w01 = ay * fz, w11 = fy * fz;          // (y, z)
    gx = (u[1] - u[0]) * w00 + (u[3] - u[2]) * w10 + (u[5] - u[4]) * w01 + (u[7] - u[6]) * w11;
    const float x00 = ax * az, x10 = fx * az, x01 = ax * fz, x11 = fx * fz;          // (x, z)
    gy = (u[2] - u[0]) * x00 + (u[3] - u[1]) * x10 + (u[6] - u[4]) * x01 + (u[7] - u[5]) * x11;
    const float y00 = ax * ay, y10 = fx * ay, y01 = ax * fy, y11 = fx * fy;          // (x, y)
    gz = (u[4] - u[0]) * y00 + (u[5] - u[1]) * y10 + (u[6] - u[2]) * y01 + (u[7] - u[3]) * y11;
}

// what the BA-mode walks need besides the planes
struct RayGradArgs {
    const float* hash_params; double bl[3]; float* g_o; float* g_d;
};
__device__ __forceinline__ void raygrad_flush(const RayGradArgs& rg, long long r, const float (&so)[3], const float (&sd)[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {                      // through the float64 normalisation (:388)
        if (rg.g_o && so[a] != 0.f) atomicAdd(rg.g_o + 3 * r + a, (float)((double)so[a] / rg.bl[a]));
        if (rg.g_d && sd[a] != 0.f) atomicAdd(rg.g_d + 3 * r + a, (float)((double)sd[a] / rg.bl[a]));
    }
}

struct ScatterRep {
    unsigned k[RF_MAX_LEVELS];            // replicas per level (power of two; 1 = accumulate straight into g_hash)
    unsigned base[RF_MAX_LEVELS];         // first entry of the level's replica block in the scratch (float2 units)
};

// BA: the same walk also differentiates the trilinear weights of its level against the corner VALUES (gathered once per
// cell run) and sums dL/d xn and z * dL/d xn over the ray's samples — the hash-level part of raygrad_walk_kernel, without
// reading the planes a second time.
template <bool BA>
__global__ void __launch_bounds__(128) scatter_walk_kernel(GridDev hg, ScatterRep rep, const float* __restrict__ xn, const float* __restrict__ dfeat,
                                                           long long P, long long N, int S, int seg, int level0, float* __restrict__ g_hash,
                                                           float* __restrict__ g_rep, RayGradArgs rg) {
    const int l = blockIdx.y + level0;
    const long long unit = blockIdx.x * 128ll + threadIdx.x;                // (ray, segment), as in encode_walk_kernel
    const long long r = unit % N;
    const int s_begin = (int)(unit / N) * seg;
    if (s_begin >= S) return;
    const int s_end = min(S, s_begin + seg);
    const long long first = (long long)s_begin * N + r;
    const unsigned size = hg.size[l];
    float2* gtab = (rep.k[l] > 1) ? reinterpret_cast<float2*>(g_rep) + rep.base[l] + (size_t)(blockIdx.x & (rep.k[l] - 1)) * size
                                  : reinterpret_cast<float2*>(g_hash) + hg.offset[l];
    const float* xs = xn + first; const float* ys = xn + P + first; const float* zs = xn + 2 * P + first;
    const float2* dj = reinterpret_cast<const float2*>(dfeat) + (long long)l * P + first;
    const float scale = hg.scale[l];
    CornerIndexer ci; ci.init(hg.is_hash != 0, size, hg.res[l]);
    unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false, nz = false;
    float2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
    float xa = __ldg(xs), ya = __ldg(ys), za = __ldg(zs); float2 da = __ldg(dj);
    const float* ts = xn + 3 * P + first;
    float ta = BA ? __ldg(ts) : 0.f;
    const float2* tab = BA ? reinterpret_cast<const float2*>(rg.hash_params) + hg.offset[l] : nullptr;
    float2 v[BA ? 8 : 1];
    bool vhave = false;                                                      // v holds the corners of cell (pcx, pcy, pcz)
    float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
    for (int s = s_begin; s <= s_end; ++s) {
        unsigned cx = 0, cy = 0, cz = 0; float fx = 0.f, fy = 0.f, fz = 0.f;
        const float2 d = da; const float t = ta;
        const bool last = (s == s_end);
        if (!last) {
            pos_fract(xa, scale, cx, fx); pos_fract(ya, scale, cy, fy); pos_fract(za, scale, cz, fz);
            if (s + 1 < s_end) {
                xs += N; ys += N; zs += N; dj += N; xa = __ldg(xs); ya = __ldg(ys); za = __ldg(zs); da = __ldg(dj);
                if (BA) { ts += N; ta = __ldg(ts); }
            }
        }
        if (have && (last || cx != pcx || cy != pcy || cz != pcz)) {
            if (nz) {
                unsigned idx[8];
                ci.cell(pcx, pcy, pcz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) atomicAdd(gtab + idx[c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
            nz = false; vhave = false;
        }
        if (last) break;
        pcx = cx; pcy = cy; pcz = cz; have = true;
        const bool dnz = d.x != 0.f || d.y != 0.f;
        nz = nz || dnz;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float w = corner_weight(c, fx, fy, fz);
            acc[c].x = fmaf(w, d.x, acc[c].x); acc[c].y = fmaf(w, d.y, acc[c].y);
        }
        if (BA && dnz) {
            if (!vhave) {
                unsigned idx[8];
                ci.cell(cx, cy, cz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + idx[c]);
                vhave = true;
            }
            float u[8], gx, gy, gz;
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = fmaf(d.x, v[c].x, d.y * v[c].y);
            tri_grad(u, fx, fy, fz, gx, gy, gz);
            const float dx0 = gx * scale, dx1 = gy * scale, dx2 = gz * scale;
            so[0] += dx0; so[1] += dx1; so[2] += dx2;
            sd[0] = fmaf(t, dx0, sd[0]); sd[1] = fmaf(t, dx1, sd[1]); sd[2] = fmaf(t, dx2, sd[2]);
        }
    }
    if (BA) raygrad_flush(rg, r, so, sd);
}


// Ray gradients (BA mode).  Thread = (ray[, segment], level) as in the other walks: level < L differentiates the
// trilinear weights of one hash level against the feature gradients dfeat (corner VALUES cached per cell run, Appendix
// B6); level L does the same for the GBV texel gradient dgb and adds the OneBlob part dxb the decoder backward wrote.
// Each thread sums d xn and z * d xn over its samples and finishes with six reductions onto the ray's rows:
//   dL/d rays_o = sum_s dL/d pts,  dL/d rays_d = sum_s z_s dL/d pts,  pts = o + d z,  d pts = d xn / (b1 - b0)  (:443, :388)
__global__ void __launch_bounds__(128) raygrad_walk_kernel(GridDev hg, GridDev gg, const float* __restrict__ hash_params,
                                                           const float* __restrict__ gbv_params, const float* __restrict__ xn,
                                                           const float* __restrict__ dfeat, const float* __restrict__ dgb,
                                                           const float* __restrict__ dxb, long long P, long long N, int S, int seg,
                                                           int level0, RayGradArgs rg) {
    const int l = blockIdx.y + level0, L = hg.n_levels;
    const long long unit = blockIdx.x * 128ll + threadIdx.x;
    const long long r = unit % N;
    const int s_begin = (int)(unit / N) * seg;
    if (s_begin >= S) return;
    const int s_end = min(S, s_begin + seg);
    const long long first = (long long)s_begin * N + r;
    const float* xs = xn + first; const float* ys = xn + P + first; const float* zs = xn + 2 * P + first; const float* ts = xn + 3 * P + first;
    unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false;
    CornerIndexer ci;
    unsigned idx[8];
    float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
    float xa = __ldg(xs), ya = __ldg(ys), za = __ldg(zs), ta = __ldg(ts);
    if (l < L) {
        const float scale = hg.scale[l];
        ci.init(hg.is_hash != 0, hg.size[l], hg.res[l]);
        const float2* tab = reinterpret_cast<const float2*>(hash_params) + hg.offset[l];
        const float2* dj = reinterpret_cast<const float2*>(dfeat) + (long long)l * P + first;
        float2 v[8];
        float2 da = __ldg(dj);
        for (int s = s_begin; s < s_end; ++s) {
            const float x = xa, y = ya, z = za, t = ta; const float2 d = da;
            if (s + 1 < s_end) { xs += N; ys += N; zs += N; ts += N; dj += N; xa = __ldg(xs); ya = __ldg(ys); za = __ldg(zs); ta = __ldg(ts); da = __ldg(dj); }
            if (d.x == 0.f && d.y == 0.f) continue;                      // masked sample: no gradient
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
                ci.cell(cx, cy, cz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + idx[c]);
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float u[8], gx, gy, gz;
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = fmaf(d.x, v[c].x, d.y * v[c].y);
            tri_grad(u, fx, fy, fz, gx, gy, gz);
            const float dx0 = gx * scale, dx1 = gy * scale, dx2 = gz * scale;
            so[0] += dx0; so[1] += dx1; so[2] += dx2;
            sd[0] = fmaf(t, dx0, sd[0]); sd[1] = fmaf(t, dx1, sd[1]); sd[2] = fmaf(t, dx2, sd[2]);
        }
    } else {
        const float scale = gg.scale[0];
        ci.init(false, gg.size[0], gg.res[0]);
        const float4* tab = reinterpret_cast<const float4*>(gbv_params);
        const float4* dj = reinterpret_cast<const float4*>(dgb) + first;
        const float* b0 = dxb + first; const float* b1 = dxb + P + first; const float* b2 = dxb + 2 * P + first;
        float4 v[8];
        float4 da = __ldg(dj); float ba0 = __ldg(b0), ba1 = __ldg(b1), ba2 = __ldg(b2);
        for (int s = s_begin; s < s_end; ++s) {
            const float x = xa, y = ya, z = za, t = ta; const float4 d = da; const float e0 = ba0, e1 = ba1, e2 = ba2;
            if (s + 1 < s_end) {
                xs += N; ys += N; zs += N; ts += N; dj += N; b0 += N; b1 += N; b2 += N;
                xa = __ldg(xs); ya = __ldg(ys); za = __ldg(zs); ta = __ldg(ts); da = __ldg(dj); ba0 = __ldg(b0); ba1 = __ldg(b1); ba2 = __ldg(b2);
            }
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
                ci.cell(cx, cy, cz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + idx[c]);
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float u[8], gx, gy, gz;
#pragma unroll
            for (int c = 0; c < 8; ++c) u[c] = d.x * v[c].x + d.y * v[c].y + d.z * v[c].z + d.w * v[c].w;
            tri_grad(u, fx, fy, fz, gx, gy, gz);
            const float dx0 = fmaf(gx, scale, e0), dx1 = fmaf(gy, scale, e1), dx2 = fmaf(gz, scale, e2);
            so[0] += dx0; so[1] += dx1; so[2] += dx2;
            sd[0] = fmaf(t, dx0, sd[0]); sd[1] = fmaf(t, dx1, sd[1]); sd[2] = fmaf(t, dx2, sd[2]);
        }
    }
    raygrad_flush(rg, r, so, sd);
}

// g_hash[level entries] += sum over the level's replicas.  blockIdx.y = level.
__global__ void __launch_bounds__(256) replica_reduce_kernel(GridDev hg, ScatterRep rep, const float* __restrict__ g_rep, float* __restrict__ g_hash) {
    const int l = blockIdx.y;
    const unsigned K = rep.k[l], size = hg.size[l];
    if (K <= 1) return;
    const float2* src = reinterpret_cast<const float2*>(g_rep) + rep.base[l];
    float2* dst = reinterpret_cast<float2*>(g_hash) + hg.offset[l];
    for (unsigned e = blockIdx.x * 256u + threadIdx.x; e < size; e += gridDim.x * 256u) {
        float2 a = make_float2(0.f, 0.f);
        for (unsigned k = 0; k < K; ++k) { float2 v = __ldg(src + (size_t)k * size + e); a.x += v.x; a.y += v.y; }
        float2 o = dst[e]; o.x += a.x; o.y += a.y; dst[e] = o;
    }
}

// Replication plan: K = 2^20 / size rounded up to a power of two, clamped to [1, 32].  Returns the scratch entries used.
static size_t scatter_plan(const GridDev& hg, long long n_rays, ScatterRep& rep) {
    static int budget0 = -1;
    if (budget0 < 0) { const char* e = getenv("RF_SCATTER_REPLICA_ENTRIES"); budget0 = e ? atoi(e) : (1 << 20); }
    // contention grows with the batch: small batches (a few thousand rays, the reference's training batches) get few or
    // no replicas, so that zeroing and folding them does not become their fixed cost
    const long long budget = std::min<long long>(budget0, 16 * n_rays);
    size_t total = 0;
    for (int l = 0; l < RF_MAX_LEVELS; ++l) {
        rep.k[l] = 1; rep.base[l] = 0;
        if (l >= hg.n_levels) continue;
        unsigned k = 1;
        while (k < 32 && (size_t)k * hg.size[l] < (size_t)budget) k <<= 1;
        if (budget <= 0) k = 1;
        rep.k[l] = k;
        if (k > 1) { rep.base[l] = (unsigned)total; total += (size_t)k * hg.size[l]; }
    }
    return total;
}
size_t scatter_scratch_floats(const GridDev& hg, long long n_rays) { ScatterRep rep; return 2 * scatter_plan(hg, n_rays, rep); }

// samples per walking thread: the whole ray for big batches; for small ones segments of >= 8 samples such that a level
// launch has ~32 k threads
static int walk_segment(long long n_rays, int S) {
    long long nseg = std::min<long long>((32768 + n_rays - 1) / std::max<long long>(n_rays, 1), std::max(1, S / 8));
    if (nseg < 1) nseg = 1;
    return (int)((S + nseg - 1) / nseg);
}

// point queries: n "rays" of one sample each (planes degenerate to [n]; the walk is one step long)
int launch_encode_points(const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* x, long long n, float* feat, cudaStream_t s) {
    const int L = hg.n_levels;
    float* xn = feat + (2ll * L + 4) * n;
    point_pos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, xn);
    RF_CHECK_LAUNCH("point_pos_kernel");
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)(L + 1));
    ProfScope ps(RF_PROF_ENCODE, s);
    encode_walk_kernel<<<grid, 128, 0, s>>>(hg, gg, p->hash_params, p->gbv_params, xn, n, n, 1, 1, 0, feat);
    RF_CHECK_LAUNCH("encode_walk_kernel");
    return 0;
}

int launch_encode(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o, const float* rays_d,
                  const float* z_vals, long long P, float* feat, cudaStream_t s) {
    const int L = hg.n_levels;
    float* xn = feat + (2ll * L + 4) * P;
    {
        static bool attr_done = false;
        if (!attr_done) { cudaFuncSetAttribute(ray_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kMaxS * 33 * (int)sizeof(float)); attr_done = true; }
        ProfScope ps(RF_PROF_RAY_POS, s);
        ray_pos_kernel<<<(unsigned)((k.n_rays + 31) / 32), 256, 4 * k.S * 33 * sizeof(float), s>>>(k, rays_o, rays_d, z_vals, P, xn);
    }
    RF_CHECK_LAUNCH("ray_pos_kernel");
    const int seg = walk_segment(k.n_rays, k.S);
    const long long units = k.n_rays * ((k.S + seg - 1) / seg);
    dim3 grid((unsigned)((units + 127) / 128), (unsigned)(L + 1));
    if (prof_enabled() && getenv("RF_DEBUG_PER_LEVEL")) {                 // per-level timing (diagnostics only)
        for (int l = 0; l <= L; ++l) {
            ProfScope pl(RF_PROF_ENCODE_LEVEL0 + l, s);
            encode_walk_kernel<<<dim3(grid.x, 1), 128, 0, s>>>(hg, gg, p->hash_params, p->gbv_params, xn, P, k.n_rays, k.S, seg, l, feat);
        }
        RF_CHECK_LAUNCH("encode_walk_kernel");
        return 0;
    }
    ProfScope ps(RF_PROF_ENCODE, s);
    encode_walk_kernel<<<grid, 128, 0, s>>>(hg, gg, p->hash_params, p->gbv_params, xn, P, k.n_rays, k.S, seg, 0, feat);
    RF_CHECK_LAUNCH("encode_walk_kernel");
    return 0;
}

// g_rep: scatter_scratch_floats() floats of scratch for the replicas (zeroed here).  rg (BA mode, optional): the walk
// also accumulates the hash-level part of the ray gradients (g_o / g_d must be zeroed by the caller).
int launch_scatter(const RayK& k, const GridDev& hg, long long P, const float* feat, const float* dfeat, float* g_hash, float* g_rep,
                   const RayGradArgs* rg, cudaStream_t s) {
    const int L = hg.n_levels;
    const float* xn = feat + (2ll * L + 4) * P;
    ScatterRep rep;
    const size_t rep_entries = scatter_plan(hg, k.n_rays, rep);
    if (rep_entries) {
        cudaError_t e = cudaMemsetAsync(g_rep, 0, rep_entries * sizeof(float2), s);
        if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(scatter replicas): %s", cudaGetErrorString(e));
    }
    const int seg = walk_segment(k.n_rays, k.S);
    const long long units = k.n_rays * ((k.S + seg - 1) / seg);
    const unsigned gx = (unsigned)((units + 127) / 128);
    RayGradArgs none{};
    if (prof_enabled() && getenv("RF_DEBUG_PER_LEVEL") && !rg) {          // per-level timing (diagnostics only)
        for (int l = 0; l < L; ++l) {
            ProfScope pl(RF_PROF_SCATTER_LEVEL0 + l, s);
            scatter_walk_kernel<false><<<dim3(gx, 1), 128, 0, s>>>(hg, rep, xn, dfeat, P, k.n_rays, k.S, seg, l, g_hash, g_rep, none);
        }
    } else {
        ProfScope ps(RF_PROF_SCATTER, s);
        if (rg) scatter_walk_kernel<true><<<dim3(gx, L), 128, 0, s>>>(hg, rep, xn, dfeat, P, k.n_rays, k.S, seg, 0, g_hash, g_rep, *rg);
        else scatter_walk_kernel<false><<<dim3(gx, L), 128, 0, s>>>(hg, rep, xn, dfeat, P, k.n_rays, k.S, seg, 0, g_hash, g_rep, none);
    }
    RF_CHECK_LAUNCH("scatter_walk_kernel");
    if (rep_entries) {
        replica_reduce_kernel<<<dim3(64, L), 256, 0, s>>>(hg, rep, g_rep, g_hash);
        RF_CHECK_LAUNCH("replica_reduce_kernel");
    }
    return 0;
}

// BA mode: dfeat [L][P][2], dgb [P][4], dxb [3][P] (all sample-major planes) -> g_rays_o / g_rays_d [N][3] (overwritten).
// With a table gradient the hash levels ride on the scatter walk and only the GBV / OneBlob level runs here.
int launch_scatter_raygrad(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, long long P, const float* feat,
                           const float* dfeat, const float* dgb, const float* dxb, float* g_hash, float* g_rep, float* g_o, float* g_d,
                           cudaStream_t s) {
    const int L = hg.n_levels;
    const float* xn = feat + (2ll * L + 4) * P;
    cudaError_t e = cudaSuccess;
    if (g_o) e = cudaMemsetAsync(g_o, 0, 3 * k.n_rays * sizeof(float), s);
    if (e == cudaSuccess && g_d) e = cudaMemsetAsync(g_d, 0, 3 * k.n_rays * sizeof(float), s);
    if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(ray gradients): %s", cudaGetErrorString(e));
    RayGradArgs rg{p->hash_params, {k.bl[0], k.bl[1], k.bl[2]}, g_o, g_d};
    if (g_hash) { int rc = launch_scatter(k, hg, P, feat, dfeat, g_hash, g_rep, &rg, s); if (rc) return rc; }
    const int level0 = g_hash ? L : 0, nlev = g_hash ? 1 : L + 1;
    const int seg = walk_segment(k.n_rays, k.S);
    const long long units = k.n_rays * ((k.S + seg - 1) / seg);
    ProfScope ps(RF_PROF_RAY_GRAD, s);
    raygrad_walk_kernel<<<dim3((unsigned)((units + 127) / 128), nlev), 128, 0, s>>>(hg, gg, p->hash_params, p->gbv_params, xn, dfeat, dgb, dxb, P,
                                                                                  k.n_rays, k.S, seg, level0, rg);
    RF_CHECK_LAUNCH("raygrad_walk_kernel");
    return 0;
}

}  // namespace rf
