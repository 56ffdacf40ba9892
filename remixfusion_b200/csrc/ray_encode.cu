// ray_encode.cu — the gather / scatter half of the Stage-2 tensor-core path (mlp_precision 1).
//
//   ray_pos_kernel      pts = o + d*z, float64 normalisation by the bounding box (model/scene_rep.py:443,388) -> xn [3][P]
//   encode_walk_kernel  hash levels + GBV trilinear features (tiny-cuda-nn grid forward, SURVEY Appendix B2-B4;
//                       model/scene_rep.py:325,329) -> level-major feature planes
//   scatter_walk_kernel hash-table gradient (Appendix B5; autograd of model/scene_rep.py:325)
//
// Both grid kernels run one thread per (ray segment, level) that WALKS the samples of its segment in order: samples
// along a ray are sorted in depth (model/scene_rep.py:428), so consecutive samples stay in the same grid cell for a
// while on all but the finest levels.  The forward keeps the 8 corner values in registers until the cell changes
// (one gather per cell run instead of one per sample); the backward accumulates the 8 corner gradients in registers
// and issues its 8 vector reductions (RED.ADD.F32x2) only when the cell changes.  The per-sample arithmetic
// (pos = fma(scale, x, 0.5), corner order, fma accumulation order) is the one grid_encode.cuh uses everywhere, so the
// features are bit-identical to the thread-per-sample kernels.
//
// Workspace layout (floats), P = n_rays * S:   [0, 2L*P) hash features [L][P][2];  [2L*P, 2L*P + 4P) GBV [P][4];
//                                              then xn [3][P].
#include "ray_common.cuh"

namespace rf {

__global__ void __launch_bounds__(256) ray_pos_kernel(RayK k, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                      const float* __restrict__ z_vals, long long P, float* __restrict__ xn) {
    long long p = blockIdx.x * 256ll + threadIdx.x;
    if (p >= P) return;
    float x[3];
    sample_x(k, rays_o, rays_d, p / k.S, z_vals[p], x);
    xn[p] = x[0]; xn[P + p] = x[1]; xn[2 * P + p] = x[2];
}

// unit u -> (ray, segment): lanes of a warp are consecutive rays at the same segment index
struct WalkGeom { long long n_rays; int S, seg, nseg; };

__device__ __forceinline__ bool walk_unit(const WalkGeom& w, long long u, long long& p0, int& n) {
    if (u >= w.n_rays * w.nseg) return false;
    long long r = u % w.n_rays; int sg = (int)(u / w.n_rays);
    int s0 = sg * w.seg;
    n = min(w.seg, w.S - s0);
    p0 = r * w.S + s0;
    return n > 0;
}

// blockIdx.y = level (0..L-1 hash levels, L = GBV).
__global__ void __launch_bounds__(128) encode_walk_kernel(GridDev hg, GridDev gg, const float* __restrict__ hash_params,
                                                          const float* __restrict__ gbv_params, const float* __restrict__ xn,
                                                          long long P, WalkGeom wg, float* __restrict__ feat) {
    const int l = blockIdx.y, L = hg.n_levels;
    long long p0; int n;
    if (!walk_unit(wg, blockIdx.x * 128ll + threadIdx.x, p0, n)) return;
    const float* xs = xn + p0; const float* ys = xn + P + p0; const float* zs = xn + 2 * P + p0;
    unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false;
    if (l < L) {
        const float scale = hg.scale[l];
        const unsigned size = hg.size[l], res = hg.res[l];
        const float2* tab = reinterpret_cast<const float2*>(hash_params) + hg.offset[l];
        float2* out = reinterpret_cast<float2*>(feat) + (long long)l * P + p0;
        float2 v[8];
        for (int i = 0; i < n; ++i) {
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(xs[i], scale, cx, fx); pos_fract(ys[i], scale, cy, fy); pos_fract(zs[i], scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    v[c] = __ldg(tab + grid_index(hg.is_hash, size, res, cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)));
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float f0 = 0.f, f1 = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float w = corner_weight(c, fx, fy, fz);
                f0 = fmaf(w, v[c].x, f0); f1 = fmaf(w, v[c].y, f1);
            }
            out[i] = make_float2(f0, f1);
        }
    } else {
        const float scale = gg.scale[0];
        const unsigned size = gg.size[0], res = gg.res[0];
        const float4* tab = reinterpret_cast<const float4*>(gbv_params);
        float4* out = reinterpret_cast<float4*>(feat + 2ll * L * P) + p0;
        float4 v[8];
        for (int i = 0; i < n; ++i) {
            unsigned cx, cy, cz; float fx, fy, fz;
            pos_fract(xs[i], scale, cx, fx); pos_fract(ys[i], scale, cy, fy); pos_fract(zs[i], scale, cz, fz);
            if (!have || cx != pcx || cy != pcy || cz != pcz) {
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    v[c] = __ldg(tab + grid_index(false, size, res, cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1)));
                pcx = cx; pcy = cy; pcz = cz; have = true;
            }
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float w = corner_weight(c, fx, fy, fz);
                o.x = fmaf(w, v[c].x, o.x); o.y = fmaf(w, v[c].y, o.y); o.z = fmaf(w, v[c].z, o.z); o.w = fmaf(w, v[c].w, o.w);
            }
            out[i] = o;
        }
    }
}

// Table-gradient scatter.  dfeat [L][P][2].  blockIdx.y = level.
__global__ void __launch_bounds__(128) scatter_walk_kernel(GridDev hg, const float* __restrict__ xn, const float* __restrict__ dfeat,
                                                           long long P, WalkGeom wg, float* __restrict__ g_hash) {
    const int l = blockIdx.y;
    long long p0; int n;
    if (!walk_unit(wg, blockIdx.x * 128ll + threadIdx.x, p0, n)) return;
    const float* xs = xn + p0; const float* ys = xn + P + p0; const float* zs = xn + 2 * P + p0;
    const float2* dj = reinterpret_cast<const float2*>(dfeat) + (long long)l * P + p0;
    const float scale = hg.scale[l];
    const unsigned size = hg.size[l], res = hg.res[l];
    float2* gtab = reinterpret_cast<float2*>(g_hash) + hg.offset[l];
    unsigned pcx = 0, pcy = 0, pcz = 0; bool have = false;
    float2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = make_float2(0.f, 0.f);
    for (int i = 0; i <= n; ++i) {
        unsigned cx = 0, cy = 0, cz = 0; float fx = 0.f, fy = 0.f, fz = 0.f;
        float2 d = make_float2(0.f, 0.f);
        const bool last = (i == n);
        if (!last) {
            pos_fract(xs[i], scale, cx, fx); pos_fract(ys[i], scale, cy, fy); pos_fract(zs[i], scale, cz, fz);
            d = dj[i];
        }
        if (have && (last || cx != pcx || cy != pcy || cz != pcz)) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (acc[c].x != 0.f || acc[c].y != 0.f)
                    atomicAdd(gtab + grid_index(hg.is_hash, size, res, pcx + (c & 1), pcy + ((c >> 1) & 1), pcz + ((c >> 2) & 1)), acc[c]);
                acc[c] = make_float2(0.f, 0.f);
            }
        }
        if (last) break;
        pcx = cx; pcy = cy; pcz = cz; have = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float w = corner_weight(c, fx, fy, fz);
            acc[c].x = fmaf(w, d.x, acc[c].x); acc[c].y = fmaf(w, d.y, acc[c].y);
        }
    }
}

static WalkGeom walk_geom(long long n_rays, int S) {
    // Whole rays per thread when there are enough rays to fill the machine; otherwise cut rays into segments (odd
    // length) so that small training batches (a few thousand rays) still expose enough parallelism.
    WalkGeom w; w.n_rays = n_rays; w.S = S;
    long long want = (long long)num_sms() * 2048 / 16;      // units per level for a full machine
    int nseg = 1;
    while (n_rays * nseg < want && S / (nseg + 1) >= 7) ++nseg;
    int seg = (S + nseg - 1) / nseg;
    if (nseg > 1 && (seg & 1) == 0) ++seg;
    w.seg = seg; w.nseg = (S + seg - 1) / seg;
    return w;
}

int launch_encode(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o, const float* rays_d,
                  const float* z_vals, long long P, float* feat, cudaStream_t s) {
    const int L = hg.n_levels;
    float* xn = feat + (2ll * L + 4) * P;
    {
        ProfScope ps(RF_PROF_RAY_POS, s);
        ray_pos_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(k, rays_o, rays_d, z_vals, P, xn);
    }
    RF_CHECK_LAUNCH("ray_pos_kernel");
    WalkGeom wg = walk_geom(k.n_rays, k.S);
    long long units = wg.n_rays * wg.nseg;
    dim3 grid((unsigned)((units + 127) / 128), (unsigned)(L + 1));
    ProfScope ps(RF_PROF_ENCODE, s);
    encode_walk_kernel<<<grid, 128, 0, s>>>(hg, gg, p->hash_params, p->gbv_params, xn, P, wg, feat);
    RF_CHECK_LAUNCH("encode_walk_kernel");
    return 0;
}

int launch_scatter(const RayK& k, const GridDev& hg, long long P, const float* feat, const float* dfeat, float* g_hash, cudaStream_t s) {
    const int L = hg.n_levels;
    const float* xn = feat + (2ll * L + 4) * P;
    WalkGeom wg = walk_geom(k.n_rays, k.S);
    long long units = wg.n_rays * wg.nseg;
    dim3 grid((unsigned)((units + 127) / 128), (unsigned)L);
    ProfScope ps(RF_PROF_SCATTER, s);
    scatter_walk_kernel<<<grid, 128, 0, s>>>(hg, xn, dfeat, P, wg, g_hash);
    RF_CHECK_LAUNCH("scatter_walk_kernel");
    return 0;
}

}  // namespace rf
