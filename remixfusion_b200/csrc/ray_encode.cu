// ray_encode.cu — the gather / scatter half of the Stage-2 tensor-core path (mlp_precision 1).
//
//   encode_walk4_kernel   pts = o + d*z, float64 normalisation by the bounding box (model/scene_rep.py:443,388), hash levels +
//                         GBV trilinear features (tiny-cuda-nn grid forward; model/scene_rep.py:325,329)
//                         -> hash features as ready tcgen05 operand tiles (bf16 hi / lo), GBV features, position planes
//   scatter_walk4_kernel  hash-table gradient (Appendix B5; autograd of model/scene_rep.py:325); BA mode: + the hash-level part of
//                         dL/d rays_o, dL/d rays_d
//   raygrad_walk_kernel   BA mode (model/scene_rep.py:443 under mp_slam/mapper.py:456,484-485): dL/d rays_o, dL/d rays_d through
//                         the trilinear weights of the GBV (Appendix B6) + the OneBlob part
//
// The grid kernels WALK the samples of a ray in order.  Samples along a ray are sorted in depth (model/scene_rep.py:428), so
// consecutive samples stay in the same grid cell for a while on all but the finest levels: the forward keeps the 8 corner
// values in registers until the cell changes (one gather per cell run instead of one per sample); the backward accumulates
// the 8 corner gradients in registers and issues its 8 vector reductions (RED.ADD.F32x2) only when the cell changes.
//
// Work split: a block owns 32 rays (lanes = consecutive rays, so every plane access of a warp is one contiguous segment)
// and its warps are ROLES: warp c walks hash levels c, c+4, c+8, c+12 (one 8-column operand chunk in the X-order of
// ray_common.cuh: a coarse, two middle and a fine level each, so the roles of a block finish together), the last warp walks the GBV.
// The block first computes the positions of its rays' samples ONCE into shared memory (the forward from rays_o / rays_d /
// z_vals, so no separate position pass and no position planes are re-read per level; the backward from the planes the
// forward wrote, one coalesced read), then every role reads them from there.  The per-sample arithmetic (pos = fma(scale, x,
// 0.5), corner order, fma accumulation order) is the one grid_encode.cuh uses everywhere, so features are bit-identical to
// the thread-per-sample fp32 kernels of ray_query.cu.
//
// The backward walks stop at n_live[r]: composite_bwd_kernel records, per ray, the last sample with a non-zero upstream
// gradient — samples beyond the truncation band behind the surface carry neither a rendering weight nor a loss term
// (model/scene_rep.py:124, model/utils.py:170-198), their gradient is exactly zero and nothing is computed for them.
#include <stdlib.h>
#include <algorithm>
#include "ray_common.cuh"
#include "umma.cuh"

namespace rf {

// samples per walking thread: the whole ray for big batches; for small ones segments of >= 8 samples such that a launch has
// ~32 k (ray, segment) units
static int walk_segment(long long n_rays, int S) {
    long long nseg = std::min<long long>((32768 + n_rays - 1) / std::max<long long>(n_rays, 1), std::max(1, S / 8));
    if (nseg < 1) nseg = 1;
    return (int)((S + nseg - 1) / nseg);
}

constexpr int kEncThreads = 160;        // 4 hash-quad roles + 1 GBV role, 32 (ray, segment) units per block

// Forward.  ws: the workspace of ray_common.cuh.  POINTS: z_vals holds n already-normalised positions [n][3] (point queries,
// model/scene_rep.py:212-310): n "rays" of one sample.
template <bool POINTS, int MINB = 3>
__global__ void __launch_bounds__(kEncThreads, MINB) encode_walk4_kernel(const __grid_constant__ RayK k, const __grid_constant__ GridDev hg,
                                                                   const __grid_constant__ GridDev gg, const float* __restrict__ hash_params,
                                                                   const float* __restrict__ gbv_params, const float* __restrict__ rays_o,
                                                                   const float* __restrict__ rays_d, const float* __restrict__ z_vals,
                                                                   long long P, int seg, float* __restrict__ ws) {
    extern __shared__ float sx[];                       // [4][seg][33]: normalised x, y, z and the depth along the ray
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const long long N = k.n_rays;
    const int S = k.S;
    const long long unit0 = blockIdx.x * 32ll;
    const long long n_units = N * ((S + seg - 1) / seg);
    for (int i = threadIdx.x; i < 32 * seg; i += kEncThreads) {
        const int ul = i / seg, sl = i - ul * seg;
        const long long unit = unit0 + ul;
        if (unit >= n_units) continue;
        const long long r = unit % N;
        const int s = (int)(unit / N) * seg + sl;
        if (s >= S) continue;
        float x[3], zv = 0.f;
        if (POINTS) { x[0] = z_vals[3 * r]; x[1] = z_vals[3 * r + 1]; x[2] = z_vals[3 * r + 2]; }
        else { zv = z_vals[r * S + s]; sample_x(k, rays_o, rays_d, r, zv, x); }
        sx[sl * 33 + ul] = x[0]; sx[(seg + sl) * 33 + ul] = x[1]; sx[(2 * seg + sl) * 33 + ul] = x[2]; sx[(3 * seg + sl) * 33 + ul] = zv;
    }
    __syncthreads();
    const long long unit = unit0 + lane;
    if (unit >= n_units) return;
    const long long r = unit % N;
    const int s_begin = (int)(unit / N) * seg;
    const int n_steps = min(S, s_begin + seg) - s_begin;
    long long q = (long long)s_begin * N + r;
    const float* px = sx + lane; const float* py = sx + seg * 33 + lane; const float* pz = sx + 2 * seg * 33 + lane;
    if (role < 4) {
        // hash levels role, role + 4, role + 8, role + 12 -> chunk `role` of the tile's operand block
        unsigned char* hop = reinterpret_cast<unsigned char*>(ws) + role * 2048;
        float2 v[4][8];
        unsigned pc[4][3];
        unsigned have = 0;
        for (int sl = 0; sl < n_steps; ++sl, q += N) {
            const float x = px[sl * 33], y = py[sl * 33], z = pz[sl * 33];
            // phase 1: cells of the four levels, gathers where a cell changed (four independent chains in flight)
            float fr[4][3];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int l = role + 4 * j;
                const float scale = hg.scale[l];
                unsigned cx, cy, cz;
                pos_fract(x, scale, cx, fr[j][0]); pos_fract(y, scale, cy, fr[j][1]); pos_fract(z, scale, cz, fr[j][2]);
                if (!((have >> j) & 1u) || cx != pc[j][0] || cy != pc[j][1] || cz != pc[j][2]) {
                    unsigned idx[8];
                    cell_indices(hg, l, cx, cy, cz, idx);
                    const float2* tab = reinterpret_cast<const float2*>(hash_params) + hg.offset[l];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[j][c] = __ldg(tab + idx[c]);
                    pc[j][0] = cx; pc[j][1] = cy; pc[j][2] = cz; have |= 1u << j;
                }
            }
            // phase 2: trilinear interpolation (Appendix B4 weight and accumulation order)
            float f[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f0 = 0.f, f1 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float w = corner_weight(c, fr[j][0], fr[j][1], fr[j][2]);
                    f0 = fmaf(w, v[j][c].x, f0); f1 = fmaf(w, v[j][c].y, f1);
                }
                f[2 * j] = f0; f[2 * j + 1] = f1;
            }
            uint4 hi, lo;
            umma::split8(f, hi, lo);
            unsigned char* row = hop + (q >> 7) * kHopTileBytes + (q & 127) * 16;
            *reinterpret_cast<uint4*>(row) = hi;
            *reinterpret_cast<uint4*>(row + kHopTileBytes / 2) = lo;
        }
    } else {
        // GBV level; this role also writes the position planes the decoder (OneBlob) and the backward walks read
        const float scale = gg.scale[0];
        const float4* tab = reinterpret_cast<const float4*>(gbv_params);
        float4* out = reinterpret_cast<float4*>(ws + ws_off_gbv(P));
        float* xn = ws + ws_off_xn(P);
        const float* pt = sx + 3 * seg * 33 + lane;
        float4 v[8];
        unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false;
        for (int sl = 0; sl < n_steps; ++sl, q += N) {
            const float x = px[sl * 33], y = py[sl * 33], z = pz[sl * 33];
            xn[q] = x; xn[P + q] = y; xn[2 * P + q] = z;
            if (!POINTS) xn[3 * P + q] = pt[sl * 33];
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
                unsigned idx[8];
                cell_indices(gg, 0, cx, cy, cz, idx);
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __ldg(tab + idx[c]);
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = corner_weight(c, fx, fy, fz);
                o4.x = fmaf(w, v[c].x, o4.x); o4.y = fmaf(w, v[c].y, o4.y); o4.z = fmaf(w, v[c].z, o4.z); o4.w = fmaf(w, v[c].w, o4.w);
            }
            out[q] = o4;
        }
    }
}

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

// Table-gradient scatter, run-length reduced.  dfeat [4 chunks][P][8] ; a block owns 32 (ray, segment) units, warp = role =
// LPT levels of one chunk (4 in mapping mode; 2 in BA mode, where a thread also keeps the corner VALUES of its cells to
// differentiate the trilinear weights).  Each lane accumulates the 8 corner gradients of its current cell in registers and
// issues the 8 vector reductions when the cell changes; runs whose gradients are all zero issue none.
// Small (coarse) levels are the contended ones: every ray of the batch lands on the same few thousand entries, and
// reductions onto one L2 line serialise.  Those levels accumulate into K private replicas of their gradient table
// (replica = block index mod K, so that neighbouring blocks — neighbouring pixels — never share one) which
// replica_reduce_kernel folds into the caller's table afterwards.
template <int LPT, bool BA>
__global__ void __launch_bounds__(512 / LPT, LPT == 4 ? 4 : 2) scatter_walk4_kernel(const __grid_constant__ GridDev hg, const __grid_constant__ ScatterRep rep,
                                                                  const float* __restrict__ xn,
                                                                  const float* __restrict__ dfeat, const int* __restrict__ n_live,
                                                                  long long P, long long N, int S, int seg, float* __restrict__ g_hash,
                                                                  float* __restrict__ g_rep, RayGradArgs rg, int pair16) {
    extern __shared__ float sx[];                       // [3 (+1 BA)][seg][32]
    constexpr int NT = 512 / LPT;
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const long long unit0 = blockIdx.x * 32ll;
    const long long n_units = N * ((S + seg - 1) / seg);
    // this lane's unit (every warp of the block sees the same 32 units)
    const long long unit = unit0 + lane;
    long long r = 0; int s_begin = 0, n_steps = 0;
    if (unit < n_units) {
        r = unit % N;
        s_begin = (int)(unit / N) * seg;
        const int lim = n_live ? min(S, __ldg(n_live + r)) : S;
        n_steps = max(0, min(lim, s_begin + seg) - s_begin);
    }
    // positions of the live samples: warp w loads steps w, w + NT/32, ... (lanes = consecutive rays: coalesced)
    for (int sl = role; sl < seg; sl += NT / 32) {
        if (sl < n_steps) {
            const long long q = (long long)(s_begin + sl) * N + r;
            sx[sl * 32 + lane] = __ldg(xn + q); sx[(seg + sl) * 32 + lane] = __ldg(xn + P + q); sx[(2 * seg + sl) * 32 + lane] = __ldg(xn + 2 * P + q);
            if (BA) sx[(3 * seg + sl) * 32 + lane] = __ldg(xn + 3 * P + q);
        }
    }
    __syncthreads();
    if (n_steps == 0) return;
    long long q = (long long)s_begin * N + r;
    const float* px = sx + lane; const float* py = sx + seg * 32 + lane; const float* pz = sx + 2 * seg * 32 + lane; const float* pt = sx + 3 * seg * 32 + lane;
    // this role's levels: chunk ch of the X-order, its entries j0 .. j0 + LPT - 1, i.e. levels ch + 4 (j0 + j)
    constexpr int RPC = 4 / LPT;                              // roles per chunk
    const int ch = role / RPC, j0 = (role % RPC) * LPT;
    const float* dq = dfeat + ((long long)ch * P) * 8 + 2 * j0;
    float2 acc[LPT][8];
    float2 v[BA ? LPT : 1][8];
    unsigned pc[LPT][3];
    unsigned have = 0, nz = 0, vhave = 0;
#pragma unroll
    for (int j = 0; j < LPT; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = make_float2(0.f, 0.f);
    float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
    // feature gradients are streamed from HBM: the loads of step sl + 1 are issued before step sl is processed
    float4 na, nb = make_float4(0.f, 0.f, 0.f, 0.f);
    na = __ldg(reinterpret_cast<const float4*>(dq + q * 8));
    if (LPT == 4) nb = __ldg(reinterpret_cast<const float4*>(dq + q * 8) + 1);
    for (int sl = 0; sl <= n_steps; ++sl, q += N) {
        const bool last = (sl == n_steps);
        float x = 0.f, y = 0.f, z = 0.f, t = 0.f;
        float d[2 * LPT];
        d[0] = na.x; d[1] = na.y; d[2] = na.z; d[3] = na.w;
        if (LPT == 4) { d[4] = nb.x; d[5] = nb.y; d[6] = nb.z; d[7] = nb.w; }
        if (!last) {
            x = px[sl * 32]; y = py[sl * 32]; z = pz[sl * 32];
            if (BA) t = pt[sl * 32];
            if (sl + 1 < n_steps) {
                na = __ldg(reinterpret_cast<const float4*>(dq + (q + N) * 8));
                if (LPT == 4) nb = __ldg(reinterpret_cast<const float4*>(dq + (q + N) * 8) + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < LPT; ++j) {
            const int l = ch + 4 * (j0 + j);
            const float scale = hg.scale[l];
            unsigned cx = 0, cy = 0, cz = 0; float fx = 0.f, fy = 0.f, fz = 0.f;
            if (!last) { pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz); }
            const bool hv = (have >> j) & 1u;
            if (hv && (last || cx != pc[j][0] || cy != pc[j][1] || cz != pc[j][2])) {
                if ((nz >> j) & 1u) {
                    const unsigned size = hg.size[l];
                    float2* gtab = (rep.k[l] > 1) ? reinterpret_cast<float2*>(g_rep) + rep.base[l] + (size_t)(blockIdx.x & (rep.k[l] - 1)) * size
                                                  : reinterpret_cast<float2*>(g_hash) + hg.offset[l];
                    unsigned idx[8];
                    cell_indices(hg, l, pc[j][0], pc[j][1], pc[j][2], idx);
                    // The two corners of an x edge are neighbours in memory whenever the low corner's entry index is even: a dense
                    // level has stride 1 in x, and on a hashed level x enters the index as x ^ (...) so x -> x + 1 flips only bit 0
                    // for even x.  Such a pair is ONE 16-byte vector reduction (red.global.add.v4.f32) into one 32-byte sector
                    // instead of two 8-byte ones: the kernel is bound by the L2 atomic unit, which counts sector operations.
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const unsigned i0 = idx[c], i1 = idx[c + 1];
                        if (!BA && pair16 && (i0 ^ i1) == 1u) {           // (BA variant: register-bound, the extra selects cost more than they save: 9.1 vs 8.9 ms)
                            const float2 a = (i0 < i1) ? acc[j][c] : acc[j][c + 1], b = (i0 < i1) ? acc[j][c + 1] : acc[j][c];
                            atomicAdd(reinterpret_cast<float4*>(gtab + (i0 & ~1u)), make_float4(a.x, a.y, b.x, b.y));
                        } else {
                            atomicAdd(gtab + i0, acc[j][c]); atomicAdd(gtab + i1, acc[j][c + 1]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[j][c] = make_float2(0.f, 0.f);
                nz &= ~(1u << j); vhave &= ~(1u << j);
            }
            if (last) continue;
            pc[j][0] = cx; pc[j][1] = cy; pc[j][2] = cz; have |= 1u << j;
            const float d0 = d[2 * j], d1 = d[2 * j + 1];
            const bool dnz = d0 != 0.f || d1 != 0.f;
            if (dnz) nz |= 1u << j;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float w = corner_weight(c, fx, fy, fz);
                acc[j][c].x = fmaf(w, d0, acc[j][c].x); acc[j][c].y = fmaf(w, d1, acc[j][c].y);
            }
            if (BA && dnz) {
                if (!((vhave >> j) & 1u)) {
                    unsigned idx[8];
                    cell_indices(hg, l, cx, cy, cz, idx);
                    const float2* tab = reinterpret_cast<const float2*>(rg.hash_params) + hg.offset[l];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[BA ? j : 0][c] = __ldg(tab + idx[c]);
                    vhave |= 1u << j;
                }
                float u[8], gx, gy, gz;
#pragma unroll
                for (int c = 0; c < 8; ++c) u[c] = fmaf(d0, v[BA ? j : 0][c].x, d1 * v[BA ? j : 0][c].y);
                tri_grad(u, fx, fy, fz, gx, gy, gz);
                const float dx0 = gx * scale, dx1 = gy * scale, dx2 = gz * scale;
                so[0] += dx0; so[1] += dx1; so[2] += dx2;
                sd[0] = fmaf(t, dx0, sd[0]); sd[1] = fmaf(t, dx1, sd[1]); sd[2] = fmaf(t, dx2, sd[2]);
            }
        }
    }
    if (BA) raygrad_flush(rg, r, so, sd);
}

// Ray gradients (BA mode).  Thread = (ray[, segment], level): level < L differentiates the trilinear weights of one hash
// level against the feature gradients dfeat (corner VALUES cached per cell run, Appendix B6) — only launched when no table
// gradient is requested, otherwise the hash levels ride on scatter_walk4_kernel<., true>; level L does the same for the GBV
// texel gradient dgb and adds the OneBlob part dxb the decoder backward wrote.  Each thread sums d xn and z * d xn over its
// live samples and finishes with six reductions onto the ray's rows:
//   dL/d rays_o = sum_s dL/d pts,  dL/d rays_d = sum_s z_s dL/d pts,  pts = o + d z,  d pts = d xn / (b1 - b0)  (:443, :388)
__global__ void __launch_bounds__(128) raygrad_walk_kernel(GridDev hg, GridDev gg, const float* __restrict__ hash_params,
                                                           const float* __restrict__ gbv_params, const float* __restrict__ xn,
                                                           const float* __restrict__ dfeat, const float* __restrict__ dgb,
                                                           const float* __restrict__ dxb, const int* __restrict__ n_live, long long P,
                                                           long long N, int S, int seg, int level0, RayGradArgs rg) {
    const int l = blockIdx.y + level0, L = hg.n_levels;
    const long long unit = blockIdx.x * 128ll + threadIdx.x;
    const long long r = unit % N;
    const int s_begin = (int)(unit / N) * seg;
    if (s_begin >= S) return;
    const int lim = n_live ? min(S, __ldg(n_live + r)) : S;
    const int s_end = min(lim, s_begin + seg);
    if (s_end <= s_begin) return;
    const long long first = (long long)s_begin * N + r;
    const float* xs = xn + first; const float* ys = xn + P + first; const float* zs = xn + 2 * P + first; const float* ts = xn + 3 * P + first;
    unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false;
    unsigned idx[8];
    float so[3] = {0.f, 0.f, 0.f}, sd[3] = {0.f, 0.f, 0.f};
    float xa = __ldg(xs), ya = __ldg(ys), za = __ldg(zs), ta = __ldg(ts);
    if (l < L) {
        const float scale = hg.scale[l];
        const float2* tab = reinterpret_cast<const float2*>(hash_params) + hg.offset[l];
        const float* dj = dfeat + ((long long)(l & 3) * P + first) * 8 + 2 * (l >> 2);      // X-order: chunk l % 4, entry l / 4
        float2 v[8];
        float2 da = __ldg(reinterpret_cast<const float2*>(dj));
        for (int s = s_begin; s < s_end; ++s) {
            const float x = xa, y = ya, z = za, t = ta; const float2 d = da;
            if (s + 1 < s_end) { xs += N; ys += N; zs += N; ts += N; dj += 8 * N; xa = __ldg(xs); ya = __ldg(ys); za = __ldg(zs); ta = __ldg(ts); da = __ldg(reinterpret_cast<const float2*>(dj)); }
            if (d.x == 0.f && d.y == 0.f) continue;                      // masked sample: no gradient
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(x, scale, cx, fx); pos_fract(y, scale, cy, fy); pos_fract(z, scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
                cell_indices(hg, l, cx, cy, cz, idx);
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
                cell_indices(gg, 0, cx, cy, cz, idx);
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
    static const int budget0 = [] { const char* e = getenv("RF_SCATTER_REPLICA_ENTRIES"); return e ? atoi(e) : (1 << 20); }();
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

// the last tile of the operand blocks is only partly written when P is not a multiple of 128: its dead rows are read by
// the decoder's bulk copies (and enter the weight-gradient GEMMs with zero upstream gradients), so they must be finite
static int zero_last_tile(long long P, float* ws, cudaStream_t s) {
    if ((P & 127) == 0) return 0;
    cudaError_t e = cudaMemsetAsync(reinterpret_cast<unsigned char*>(ws) + (ws_tiles(P) - 1) * (size_t)kHopTileBytes, 0, kHopTileBytes, s);
    if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(last operand tile): %s", cudaGetErrorString(e));
    return 0;
}

// point queries: n "rays" of one sample each (planes degenerate to [n]; the walk is one step long)
int launch_encode_points(const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* x, long long n, float* feat, cudaStream_t s) {
    RayK k; memset(&k, 0, sizeof(k));
    k.n_rays = n; k.S = 1;
    int rc = zero_last_tile(n, feat, s); if (rc) return rc;
    ProfScope ps(RF_PROF_ENCODE, s);
    encode_walk4_kernel<true><<<(unsigned)((n + 31) / 32), kEncThreads, 4 * 33 * sizeof(float), s>>>(k, hg, gg, p->hash_params, p->gbv_params, nullptr,
                                                                                                  nullptr, x, n, 1, feat);
    RF_CHECK_LAUNCH("encode_walk4_kernel<points>");
    return 0;
}

int launch_encode(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o, const float* rays_d,
                  const float* z_vals, long long P, float* feat, cudaStream_t s) {
    int rc = zero_last_tile(P, feat, s); if (rc) return rc;
    const int seg = walk_segment(k.n_rays, k.S);
    const long long units = k.n_rays * ((k.S + seg - 1) / seg);
    const size_t sm = 4 * (size_t)seg * 33 * sizeof(float);
    // resident blocks per SM the register allocation is tuned for (RF_ENC_MINB = 3 | 4 | 5: 136 / 96 / 80 registers per thread;
    // measured on the bench frame: 5.50 / 6.12 / 9.77 ms — no spills beats occupancy, profiles/r2f_bench_minb3.json)
    static const int minb = [] { const char* e = getenv("RF_ENC_MINB"); return e ? atoi(e) : 3; }();
    auto fn = minb == 4 ? encode_walk4_kernel<false, 4> : minb == 5 ? encode_walk4_kernel<false, 5> : encode_walk4_kernel<false, 3>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * kMaxS * 33 * sizeof(float)));
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(encode_walk4_kernel): %s", cudaGetErrorString(e));
    ProfScope ps(RF_PROF_ENCODE, s);
    fn<<<(unsigned)((units + 31) / 32), kEncThreads, sm, s>>>(k, hg, gg, p->hash_params, p->gbv_params, rays_o, rays_d, z_vals, P, seg, feat);
    RF_CHECK_LAUNCH("encode_walk4_kernel");
    return 0;
}

// g_rep: scatter_scratch_floats() floats of scratch for the replicas (zeroed here).  rg (BA mode, optional): the walk
// also accumulates the hash-level part of the ray gradients (g_o / g_d must be zeroed by the caller).
int launch_scatter(const RayK& k, const GridDev& hg, long long P, const float* feat, const float* dfeat, const int* n_live, float* g_hash,
                   float* g_rep, const RayGradArgs* rg, cudaStream_t s) {
    const float* xn = feat + ws_off_xn(P);
    ScatterRep rep;
    const size_t rep_entries = scatter_plan(hg, k.n_rays, rep);
    if (rep_entries) {
        cudaError_t e = cudaMemsetAsync(g_rep, 0, rep_entries * sizeof(float2), s);
        if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(scatter replicas): %s", cudaGetErrorString(e));
    }
    const int seg = walk_segment(k.n_rays, k.S);
    const long long units = k.n_rays * ((k.S + seg - 1) / seg);
    const unsigned gx = (unsigned)((units + 31) / 32);
    RayGradArgs none{};
    // 16-byte vector reductions for neighbouring corner pairs need every level's table (caller's or replica) to start on a 16-byte
    // boundary: even float2 offsets on 16-byte aligned buffers (the ABI only promises 8 bytes for g_hash); RF_SCATTER_PAIR16=0 disables
    static const bool pair_env = [] { const char* e = getenv("RF_SCATTER_PAIR16"); return e ? atoi(e) != 0 : true; }();
    int pair16 = pair_env && ((((uintptr_t)g_hash) | ((uintptr_t)g_rep)) & 15) == 0;
    for (int l = 0; l < hg.n_levels; ++l) pair16 = pair16 && (hg.offset[l] % 2 == 0) && (hg.size[l] % 2 == 0) && (rep.base[l] % 2 == 0);
    {
        ProfScope ps(RF_PROF_SCATTER, s);
        cudaError_t e;
        if (rg) {
            const size_t sm = 4 * (size_t)seg * 32 * sizeof(float);
            e = cudaFuncSetAttribute(scatter_walk4_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * kMaxS * 32 * sizeof(float)));
            if (e == cudaSuccess) scatter_walk4_kernel<2, true><<<gx, 256, sm, s>>>(hg, rep, xn, dfeat, n_live, P, k.n_rays, k.S, seg, g_hash, g_rep, *rg, pair16);
        } else {
            const size_t sm = 3 * (size_t)seg * 32 * sizeof(float);
            e = cudaFuncSetAttribute(scatter_walk4_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * kMaxS * 32 * sizeof(float)));
            if (e == cudaSuccess) scatter_walk4_kernel<4, false><<<gx, 128, sm, s>>>(hg, rep, xn, dfeat, n_live, P, k.n_rays, k.S, seg, g_hash, g_rep, none, pair16);
        }
        if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(scatter_walk4_kernel): %s", cudaGetErrorString(e));
    }
    RF_CHECK_LAUNCH("scatter_walk4_kernel");
    if (rep_entries) {
        replica_reduce_kernel<<<dim3(64, hg.n_levels), 256, 0, s>>>(hg, rep, g_rep, g_hash);
        RF_CHECK_LAUNCH("replica_reduce_kernel");
    }
    return 0;
}

// BA mode: dfeat [4][P][8], dgb [P][4], dxb [3][P] (sample-major planes) -> g_rays_o / g_rays_d [N][3] (overwritten).
// With a table gradient the hash levels ride on the scatter walk and only the GBV / OneBlob level runs here.
int launch_scatter_raygrad(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, long long P, const float* feat,
                           const float* dfeat, const float* dgb, const float* dxb, const int* n_live, float* g_hash, float* g_rep, float* g_o,
                           float* g_d, cudaStream_t s) {
    const int L = hg.n_levels;
    const float* xn = feat + ws_off_xn(P);
    cudaError_t e = cudaSuccess;
    if (g_o) e = cudaMemsetAsync(g_o, 0, 3 * k.n_rays * sizeof(float), s);
    if (e == cudaSuccess && g_d) e = cudaMemsetAsync(g_d, 0, 3 * k.n_rays * sizeof(float), s);
    if (e != cudaSuccess) return set_error((int)e, "cudaMemsetAsync(ray gradients): %s", cudaGetErrorString(e));
    RayGradArgs rg{p->hash_params, {k.bl[0], k.bl[1], k.bl[2]}, g_o, g_d};
    if (g_hash) { int rc = launch_scatter(k, hg, P, feat, dfeat, n_live, g_hash, g_rep, &rg, s); if (rc) return rc; }
    const int level0 = g_hash ? L : 0, nlev = g_hash ? 1 : L + 1;
    const int seg = walk_segment(k.n_rays, k.S);
    const long long units = k.n_rays * ((k.S + seg - 1) / seg);
    ProfScope ps(RF_PROF_RAY_GRAD, s);
    raygrad_walk_kernel<<<dim3((unsigned)((units + 127) / 128), nlev), 128, 0, s>>>(hg, gg, p->hash_params, p->gbv_params, xn, dfeat, dgb, dxb, n_live,
                                                                                  P, k.n_rays, k.S, seg, level0, rg);
    RF_CHECK_LAUNCH("raygrad_walk_kernel");
    return 0;
}

}  // namespace rf
