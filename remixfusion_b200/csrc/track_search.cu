// track_search.cu — SURVEY §8f N2, second half: the kernels of the random-optimisation tracker that read the moving
// volume (model/ROtracker.py:141-400; host :436-470, :536-604).
//
//   track_row_sample_kernel   the per-row random offset of compute_vertex (:306-330): the reference seeds one XORWOW
//                             stream per pixel with subsequence = ROW index, so every pixel of a row draws the same
//                             numbers — one thread per row does the (expensive) curand_init here instead of one per pixel
//   track_vertex_kernel       compute_vertex (:273-344): back-projected vertex at depth + z offset, and the TSDF value that
//                             offset implies
//   track_normal_kernel       compute_normal (:346-400)
//   track_fitness_kernel      compute_tsdf_value (:144-271): for every pose candidate, sum over the sub-sampled pixels of
//                             |tsdf(nearest voxel of the transformed vertex) - expected tsdf|, and the number of hits
//
// The reference launches one thread per (candidate, pixel) that adds into two per-candidate floats with system-scope
// atomics (nondeterministic order).  Here a thread owns one candidate (its quaternion set-up is done once), the block
// stages a chunk of valid pixels (vertex rotated into the world frame once per pixel, not once per candidate) in shared
// memory, every thread accumulates in registers, and per-chunk partial sums are folded in a fixed order: the result is
// deterministic.  Per-term arithmetic is written with explicit round-to-nearest intrinsics in exactly the fused form
// nvcc 12.9 / ptxas emit for the reference kernel on sm_100a (read from its SASS: which products are rounded on their
// own and which are contracted into FFMA decides, in rare cases, which voxel a vertex rounds to); tests compare
// single-pixel sums — and through them every term — bit for bit with the literal reference kernel.
#include <curand_kernel.h>
#include "rf_common.cuh"

namespace rf {
namespace {

constexpr int kCand = 128;          // candidates (threads) per block
constexpr int kPixChunk = 128;      // pixels staged per shared-memory round

struct TrackCam { float k[9]; };
struct TrackPose { float R[9]; float T[3]; float ss[6]; };

__global__ void track_row_sample_kernel(int H, unsigned long long seed, float sample_range, float* __restrict__ row_sample) {
    const int pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= H) return;
    curandState state;
    curand_init(seed, pi, 0, &state);                                       // :307-309
    float sample = (curand_uniform(&state) * (sample_range + 1)) - sample_range;   // :318
    if (sample_range < 1) sample = (curand_uniform(&state) * 2 * sample_range) - sample_range;   // :321-324
    row_sample[pi] = sample;
}

__global__ void track_vertex_kernel(const float* __restrict__ depth, const float* __restrict__ row_sample, TrackCam cam, int H, int W,
                                    float cutdist, float trunc, float4* __restrict__ vertex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int pi = i / W, pj = i - pi * W;
    float depth_value = depth[i];
    if (depth_value > cutdist) depth_value = 0.f;                            // :292-294
    if (depth_value <= 0) { vertex[i] = make_float4(0.f, 0.f, 0.f, 0.f); return; }   // :296-303
    const float sample = row_sample[pi];
    const float z_val = sample * trunc;
    float gt_tsdf = -sample;                                                 // :326-334
    if (z_val < -1 * trunc) gt_tsdf = 1.0;
    if (z_val > 1 * trunc) gt_tsdf = 1.0;
    const float c_z = depth_value + z_val;                                   // :337-339
    const float c_x = ((float)pj - cam.k[0 * 3 + 2]) * c_z / cam.k[0];
    const float c_y = ((float)pi - cam.k[1 * 3 + 2]) * c_z / cam.k[1 * 3 + 1];
    vertex[i] = make_float4(c_x, c_y, c_z, gt_tsdf);
}

// Border pixels are left untouched, as in the reference (:353-355): the caller zero-fills the map once.
__global__ void track_normal_kernel(const float4* __restrict__ vertex, int H, int W, float* __restrict__ normal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W) return;
    const int pi = i / W, pj = i - pi * W;
    if (pi > H - 2 || pj > W - 2 || pi < 1 || pj < 1) return;
    const float4 c = vertex[i], l = vertex[i - 1], r = vertex[i + 1], u = vertex[i - W], d = vertex[i + W];
    if (c.z == 0 || l.z == 0 || r.z == 0 || u.z == 0 || d.z == 0) {          // :364-369
        normal[3 * i] = 0.f; normal[3 * i + 1] = 0.f; normal[3 * i + 2] = 0.f;
        return;
    }
    float hor_x = l.x - r.x, hor_y = l.y - r.y, hor_z = l.z - r.z;           // :371-377
    float ver_x = u.x - d.x, ver_y = u.y - d.y, ver_z = u.z - d.z;
    float normal_x = -hor_z * ver_y + hor_y * ver_z;                         // :379-385
    float normal_y = hor_z * ver_x - hor_x * ver_z;
    float normal_z = -hor_y * ver_x + hor_x * ver_y;
    float lens = sqrt(normal_x * normal_x + normal_y * normal_y + normal_z * normal_z);
    normal_x = normal_x / lens; normal_y = normal_y / lens; normal_z = normal_z / lens;
    if (normal_z > 0) { normal_x *= -1; normal_y *= -1; normal_z *= -1; }    // :387-391
    normal[3 * i] = normal_x; normal[3 * i + 1] = normal_y; normal[3 * i + 2] = normal_z;
}

// Device-resident state of the search loop (rf_track_random_optimization): 64 32-bit words.  The host fills R, T, ss, prev_ss,
// the loop kernels keep everything else; `n / level / cand_off / chunks` are the launch parameters of the NEXT fitness evaluation.
struct TrackDevState {
    float R[9], T[3];                     //  0..11  current_global_R / current_global_T
    float ss[6], prev_ss[6], spare[6];    // 12..29  search_size, previous_search_size
    float out9[9];                        // 30..38  cal_transform of the last iteration
    int count_particle, level_index, success, previous_success, iter, prev_frame_success;   // 39..44
    int n, level, cand_off, chunks;       // 45..48
    int succ_mask, pad[14];               // 49      bit i = iteration i succeeded
};
static_assert(sizeof(TrackDevState) == 64 * 4, "TrackDevState layout");
struct TrackPolicy {
    int iters, count_search, fix_level_index, iterative_scale, n_sms, H, W;
    float scaling, beta;
    int pst_n[20], pst_off[20], depth_level[20];
};

struct FitArgs {
    const float* tsdf; int dx, dy, dz; int ox, oy, oz; float voxel;
    const float4* vertex; const float* normal; int H, W;
    TrackCam cam; TrackPose pose;
    const float* cand; int n; int level, level_index, ph, pw;               // ph x pw = sub-sampled pixel grid
    int chunks, pix_per_chunk;
    const TrackDevState* st;                                                // device loop: pose, candidates and geometry come from here
};

__host__ __device__ inline int fit_chunks_of(int n, int pixels, int n_sms) {
    const int cand_blocks = (n + 128 - 1) / 128;
    int chunks = (16 * n_sms + cand_blocks - 1) / cand_blocks;
    const int cap = (pixels + 31) / 32;
    chunks = chunks < cap ? chunks : cap;
    if (chunks < 1) chunks = 1;
    return chunks < 512 ? chunks : 512;
}

// grid = (ceil(n / kCand), chunks); partial[(chunk * n + node) * 2 + {0,1}] = (sum, count) over the chunk's pixels
// DEV: launched with the worst-case geometry of the search loop; candidate count, pyramid level, pose and search size are read from
// the device state, and the pixel chunking is the one rf_track_fitness would have chosen (identical sums).
template <bool DEV>
__global__ void __launch_bounds__(kCand) track_fitness_kernel(FitArgs a, float* __restrict__ partial) {
    __shared__ float4 sv[kPixChunk];               // vertex rotated into the world frame (:211-213), gt tsdf
    __shared__ unsigned char sok[kPixChunk];       // pixel passes the validity tests (:182-203)
    __shared__ float s_pose[18];                   // DEV: R, T, ss
    if (DEV) {
        const TrackDevState* st = a.st;
        a.n = st->n; a.level = st->level; a.level_index = st->level_index; a.cand = a.cand + st->cand_off; a.chunks = st->chunks;
        a.ph = a.H / a.level; a.pw = a.W / a.level;
        const int pixels = a.ph * a.pw;
        a.pix_per_chunk = max(1, (pixels + a.chunks - 1) / a.chunks);
        if ((int)blockIdx.y >= a.chunks || (int)blockIdx.x * kCand >= a.n) return;       // block-uniform
        if (threadIdx.x < 9) s_pose[threadIdx.x] = st->R[threadIdx.x];
        else if (threadIdx.x < 12) s_pose[threadIdx.x] = st->T[threadIdx.x - 9];
        else if (threadIdx.x < 18) s_pose[threadIdx.x] = st->ss[threadIdx.x - 12];
        __syncthreads();
    }
    const int node = blockIdx.x * kCand + threadIdx.x;
    const bool live = node < a.n;
    const float* R = DEV ? s_pose : a.pose.R; const float* T = DEV ? s_pose + 9 : a.pose.T; const float* SS = DEV ? s_pose + 12 : a.pose.ss;
    // candidate set-up (:215-222)
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, q0 = 1.f;
    if (live) {
        c0 = a.cand[node * 6 + 0]; c1 = a.cand[node * 6 + 1]; c2 = a.cand[node * 6 + 2];     // t = c * ss is contracted below
        q1 = __fmul_rn(a.cand[node * 6 + 3], SS[3]);
        q2 = __fmul_rn(a.cand[node * 6 + 4], SS[4]);
        q3 = __fmul_rn(a.cand[node * 6 + 5], SS[5]);
        q0 = __fsqrt_rn(__fmaf_rn(-q3, q3, __fmaf_rn(-q2, q2, __fmaf_rn(-q1, q1, 1.0f))));   // :222
    }
    const float ss0 = SS[0], ss1 = SS[1], ss2 = SS[2];
    float sum = 0.f, cnt = 0.f;
    const int im_h = a.ph * a.level, im_w = a.pw * a.level;                 // :171-172
    const int p_begin = blockIdx.y * a.pix_per_chunk, p_end = min(a.ph * a.pw, p_begin + a.pix_per_chunk);
    for (int base = p_begin; base < p_end; base += kPixChunk) {
        __syncthreads();
        const int m = min(kPixChunk, p_end - base);
        for (int j = threadIdx.x; j < m; j += kCand) {
            const int p = base + j;
            const int pi = (p / a.pw) * a.level + a.level_index, pj = (p % a.pw) * a.level + a.level_index;   // :179-180
            bool ok = !(pi > im_h - 1 || pj > im_w - 1 || pi < 0 || pj < 0);                                 // :182
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                const int i = pi * a.W + pj;
                ok = !(a.normal[3 * i] == 0 && a.normal[3 * i + 1] == 0 && a.normal[3 * i + 2] == 0);        // :186
                if (ok) { v = a.vertex[i]; ok = !(v.x == 0 && v.y == 0 && v.z == 0); }                       // :201
            }
            sok[j] = ok ? 1 : 0;
            if (ok) {
                const float x = v.x, y = v.y, z = v.z;
                float global_x = __fmaf_rn(z, R[2], __fmaf_rn(x, R[0], __fmul_rn(y, R[1])));   // :211-213
                float global_y = __fmaf_rn(z, R[5], __fmaf_rn(x, R[3], __fmul_rn(y, R[4])));
                float global_z = __fmaf_rn(z, R[8], __fmaf_rn(x, R[6], __fmul_rn(y, R[7])));
                sv[j] = make_float4(global_x, global_y, global_z, v.w);
            }
        }
        __syncthreads();
        if (live) {
            for (int j = 0; j < m; ++j) {
                if (!sok[j]) continue;                                                 // block-uniform
                const float4 g = sv[j];
                const float global_x = g.x, global_y = g.y, global_z = g.z, gt_tsdf = g.w;
                // :224-227; S = -q_w
                const float q_z = __fmaf_rn(global_z, q0, __fmaf_rn(global_y, q1, -__fmul_rn(global_x, q2)));
                const float S   = __fmaf_rn(global_z, q3, __fmaf_rn(global_x, q1, __fmul_rn(global_y, q2)));
                const float q_y = __fmaf_rn(-global_z, q1, __fmaf_rn(global_x, q3, __fmul_rn(global_y, q0)));
                const float q_x = __fmaf_rn(global_z, q2, __fmaf_rn(global_x, q0, -__fmul_rn(global_y, q3)));
                // :229-231 (the translation c * search_size is contracted into the chain, then + T)
                const float x = __fadd_rn(__fmaf_rn(c0, ss0, __fmaf_rn(-q3, q_y, __fmaf_rn(q2, q_z, __fmaf_rn(q1, S, __fmul_rn(q_x, q0))))), T[0]);
                const float y = __fadd_rn(__fmaf_rn(c1, ss1, __fmaf_rn(q3, q_x, __fmaf_rn(q2, S, __fmaf_rn(q_y, q0, -__fmul_rn(q1, q_z))))), T[1]);
                const float z = __fadd_rn(__fmaf_rn(c2, ss2, __fmaf_rn(q3, S, __fmaf_rn(-q2, q_x, __fmaf_rn(q_z, q0, __fmul_rn(q1, q_y))))), T[2]);
                const float vcx = __fadd_rn(x, -T[0]), vcy = __fadd_rn(y, -T[1]), vcz = __fadd_rn(z, -T[2]);          // :233-235
                const float cam_x = __fmaf_rn(R[6], vcz, __fmaf_rn(R[0], vcx, __fmul_rn(R[3], vcy)));                  // :237-239
                const float cam_y = __fmaf_rn(R[7], vcz, __fmaf_rn(R[1], vcx, __fmul_rn(R[4], vcy)));
                const float cam_z = __fmaf_rn(R[8], vcz, __fmaf_rn(R[2], vcx, __fmul_rn(R[5], vcy)));
                const int pixel_x = __float2int_rz(__fadd_rn(__fadd_rn(a.cam.k[2], __fdiv_rn(__fmul_rn(cam_x, a.cam.k[0]), cam_z)), 0.5f));   // :241-242
                const int pixel_y = __float2int_rz(__fadd_rn(__fadd_rn(a.cam.k[5], __fdiv_rn(__fmul_rn(cam_y, a.cam.k[4]), cam_z)), 0.5f));
                if (pixel_x >= 0 && pixel_y >= 0 && pixel_x < a.W && pixel_y < a.H && cam_z >= 0) {          // :245
                    const int voxel_x = (int)roundf(__fdiv_rn(__fadd_rn(x, -(float)a.ox), a.voxel));                   // :246-248
                    const int voxel_y = (int)roundf(__fdiv_rn(__fadd_rn(y, -(float)a.oy), a.voxel));
                    const int voxel_z = (int)roundf(__fdiv_rn(__fadd_rn(z, -(float)a.oz), a.voxel));
                    if (voxel_x < 1 || voxel_x >= a.dx - 1 || voxel_y < 1 || voxel_y >= a.dy - 1 || voxel_z < 1 || voxel_z >= a.dz - 1) continue;   // :250
                    int index = voxel_z + voxel_y * a.dz + voxel_x * a.dy * a.dz;                            // :254
                    const float add_value = fabsf(__fadd_rn(__ldg(a.tsdf + index), -gt_tsdf));                      // :261
                    sum += add_value; cnt += 1.f;
                }
            }
        }
    }
    if (live) {
        float2* out = reinterpret_cast<float2*>(partial) + (size_t)blockIdx.y * a.n + node;
        *out = make_float2(sum, cnt);
    }
}

__global__ void track_fold_kernel(const float* __restrict__ partial, int n, int chunks, float* __restrict__ value, float* __restrict__ count,
                                  const TrackDevState* __restrict__ st) {
    if (st) { n = st->n; chunks = st->chunks; }
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    float s = 0.f, c = 0.f;
    for (int k = 0; k < chunks; ++k) {
        float2 v = reinterpret_cast<const float2*>(partial)[(size_t)k * n + node];
        s += v.x; c += v.y;
    }
    value[node] = s; count[node] = c;
}

// cal_transform (model/ROtracker.py:606-714): weighted mean of the FIRST `count_search` candidates (in index order) that
// fit better than candidate 0, weights = origin_tsdf - fit.  One block: every thread owns a contiguous range of
// candidates, a block scan of the per-range qualifier counts gives each qualifier its ordinal, partial sums are kept in
// double (the reference accumulates float32 products into Python floats) and folded in a fixed order.
// out[0] = success (0 / 1), out[1] = min_tsdf, out[2..8] = mean_transform (tx, ty, tz, qw, qx, qy, qz).
constexpr int kCalThreads = 1024;
__global__ void __launch_bounds__(kCalThreads, 1) track_cal_transform_kernel(const float* __restrict__ value, const float* __restrict__ count,
                                                                          const float* __restrict__ cand, int n, TrackPose pose, int count_search,
                                                                          float* __restrict__ out, TrackDevState* __restrict__ st) {
    float ss[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) ss[i] = pose.ss[i];
    if (st) {                                                                                // device loop: size, table, search size, output from the state
        n = st->n; cand += st->cand_off; out = st->out9;
#pragma unroll
        for (int i = 0; i < 6; ++i) ss[i] = st->ss[i];
    }
    __shared__ int s_cnt[kCalThreads];
    __shared__ double s_sum[9][32];
    const int t = threadIdx.x;
    auto fit = [&](int j) { return __fdiv_rn(value[j], __fadd_rn(count[j], 1e-6f)); };       // evaluate_tsdf :604
    const float origin = fit(0);                                                             // :622
    const int per = (n - 1 + kCalThreads - 1) / kCalThreads;
    const int j0 = 1 + t * per, j1 = min(n, j0 + per);
    int mine = 0;
    for (int j = j0; j < j1; ++j) mine += fit(j) < origin ? 1 : 0;                           // :638
    s_cnt[t] = mine;
    __syncthreads();
    for (int off = 1; off < kCalThreads; off <<= 1) {                                        // inclusive scan
        int v = (t >= off) ? s_cnt[t - off] : 0;
        __syncthreads();
        s_cnt[t] += v;
        __syncthreads();
    }
    int ordinal = s_cnt[t] - mine;                                                           // qualifiers before this range
    const int total = min(s_cnt[kCalThreads - 1], count_search);                             // :677-678
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // tx ty tz qx qy qz qw weight tsdf
    for (int j = j0; j < j1 && ordinal < count_search; ++j) {
        const float cur_fit = fit(j);
        if (!(cur_fit < origin)) continue;
        ++ordinal;
        const float weight = __fadd_rn(origin, -cur_fit);                                    // :647
        const float* c = cand + (size_t)j * 6;
        a[0] += (double)__fmul_rn(c[0], weight); a[1] += (double)__fmul_rn(c[1], weight); a[2] += (double)__fmul_rn(c[2], weight);   // :649-654
        a[3] += (double)__fmul_rn(c[3], weight); a[4] += (double)__fmul_rn(c[4], weight); a[5] += (double)__fmul_rn(c[5], weight);
        const float qx = __fmul_rn(c[3], ss[3]), qy = __fmul_rn(c[4], ss[4]), qz = __fmul_rn(c[5], ss[5]);           // :657-659
        const float qw = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(1.0f, -__fmul_rn(qx, qx)), -__fmul_rn(qy, qy)), -__fmul_rn(qz, qz)));   // :670
        a[6] += (double)__fmul_rn(qw, weight); a[7] += (double)weight; a[8] += (double)__fmul_rn(cur_fit, weight);                 // :671-673
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = a[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((t & 31) == 0) s_sum[k][t >> 5] = v;
    }
    __syncthreads();
    if (t == 0) {
        double r[9];
#pragma unroll 1
        for (int w = 0; w < kCalThreads / 32; ++w) {           // (not unrolled: 288 hoisted loads spilled 1.8 KB per thread)
#pragma unroll
            for (int k = 0; k < 9; ++k) r[k] = (w == 0 ? 0.0 : r[k]) + s_sum[k][w];
        }
        if (total <= 0) {                                                                    // :681-684
            out[0] = 0.f; out[1] = origin;
            for (int k = 2; k < 9; ++k) out[k] = 0.f;
        } else {
            const double sw = r[7];
            out[0] = 1.f;
            out[1] = (float)(r[8] / sw);                                                     // :687, :712
            out[2] = (float)((r[0] / sw) * (double)ss[0]);                              // :688-690
            out[3] = (float)((r[1] / sw) * (double)ss[1]);
            out[4] = (float)((r[2] / sw) * (double)ss[2]);
            const double qww = r[6] / sw, qxx = (r[3] / sw) * (double)ss[3], qyy = (r[4] / sw) * (double)ss[4],
                         qzz = (r[5] / sw) * (double)ss[5];                              // :691-694
            const double lens = 1.0 / sqrt(qww * qww + qxx * qxx + qyy * qyy + qzz * qzz);   // :701
            out[5] = (float)(qww * lens); out[6] = (float)(qxx * lens); out[7] = (float)(qyy * lens); out[8] = (float)(qzz * lens);
        }
    }
}

// Pixel chunks per candidate block: enough blocks for ~16 per SM (the per-pair chain of five IEEE divisions and a
// dependent gather needs many warps to hide), at least 32 pixels per chunk.
static int fit_chunks(int n, int pixels) { return fit_chunks_of(n, pixels, num_sms()); }

// The search policy between two fitness evaluations (model/ROtracker.py:757-826: the body of random_optimization's loop after
// cal_transform, and update_PST :495-531), one thread.  Scalars are combined in double and narrowed where the reference stores
// into its float32 arrays.  It ends by setting up the next evaluation: the reset of count_particle at the top of the next
// iteration (:758-759), the candidate table (:761-763 get_PST), the pyramid level (:764) and the pixel chunking.
__global__ void track_policy_kernel(TrackDevState* __restrict__ st, TrackPolicy pol) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int i = st->iter;
    const bool success = st->out9[0] > 0.5f;
    const float min_tsdf = st->out9[1];
    const float* mt = st->out9 + 2;                                    // tx ty tz qw qx qy qz
    int cp = st->count_particle;
    if (success) {                                                     // :776-787
        if (cp < 19) cp += 1;
        const float qw = mt[3], qx = mt[4], qy = mt[5], qz = mt[6];
        // float32 products and sums, then `2 *` / `1 -` in double narrowed to float32: what the reference's numpy 1.21.6 computes
        // (a float32 scalar times a Python int promotes to float64 there)
        auto pp = [](float a, float b, float c, float d) { return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d)); };
        auto pm = [](float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d)); };
        auto one_minus_2 = [](float v) { return (float)(1.0 - 2.0 * (double)v); };
        const float Ri[9] = {one_minus_2(pp(qy, qy, qz, qz)), 2.f * pm(qx, qy, qz, qw), 2.f * pp(qx, qz, qy, qw),
                             2.f * pp(qx, qy, qz, qw), one_minus_2(pp(qx, qx, qz, qz)), 2.f * pm(qy, qz, qx, qw),
                             2.f * pm(qx, qz, qy, qw), 2.f * pp(qy, qz, qx, qw), one_minus_2(pp(qx, qx, qy, qy))};
        for (int k = 0; k < 3; ++k) st->T[k] = __fadd_rn(st->T[k], mt[k]);
        float Rn[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                Rn[3 * r + c] = __fadd_rn(__fadd_rn(__fmul_rn(Ri[3 * r], st->R[c]), __fmul_rn(Ri[3 * r + 1], st->R[3 + c])), __fmul_rn(Ri[3 * r + 2], st->R[6 + c]));
        for (int k = 0; k < 9; ++k) st->R[k] = Rn[k];
        st->succ_mask |= 1 << i;
    }
    int li = pol.fix_level_index ? 1 : st->level_index + 5;            // :790-795
    li = li % pol.depth_level[cp];
    {   // update_PST (:495-531), min_scale = 1e-3, scale = scaling_coefficient
        const double ms = 1e-3;
        const double s_tx = fabs((double)mt[0]) + ms, s_ty = fabs((double)mt[1]) + ms, s_tz = fabs((double)mt[2]) + ms;
        const double s_qx = fabs((double)mt[4]) + ms, s_qy = fabs((double)mt[5]) + ms, s_qz = fabs((double)mt[6]) + ms;
        const double nrm = sqrt(s_tx * s_tx + s_ty * s_ty + s_tz * s_tz + s_qx * s_qx + s_qy * s_qy + s_qz * s_qz);
        const double k = (double)pol.scaling * (double)min_tsdf;
        st->ss[3] = (float)(k * (s_qx / nrm) + ms); st->ss[4] = (float)(k * (s_qy / nrm) + ms); st->ss[5] = (float)(k * (s_qz / nrm) + ms);
        st->ss[0] = (float)(k * (s_tx / nrm) + ms); st->ss[1] = (float)(k * (s_ty / nrm) + ms); st->ss[2] = (float)(k * (s_tz / nrm) + ms);
    }
    int prev = st->previous_success;
    if (prev && success) {                                             // :801-808
        for (int k = 0; k < 6; ++k) st->ss[k] = (float)((double)pol.beta * (double)st->ss[k] + (1.0 - (double)pol.beta) * (double)st->prev_ss[k]);
    } else if (success) {                                              // :810-819
        if (pol.iterative_scale) prev = 1;
        for (int k = 0; k < 6; ++k) st->prev_ss[k] = st->ss[k];
    }
    if (!success) prev = 0;                                            // :821-822
    if (i == 0) st->prev_frame_success = success ? 1 : 0;              // :824-830 (initialize_search_size aliases search_size: the host mirrors that)
    st->previous_success = prev; st->success = success ? 1 : 0; st->level_index = li; st->iter = i + 1;
    // next evaluation (:758-764)
    if (!success) cp = 0;
    st->count_particle = cp;
    st->n = pol.pst_n[cp]; st->cand_off = pol.pst_off[cp]; st->level = pol.depth_level[cp];
    st->chunks = fit_chunks_of(st->n, (pol.H / st->level) * (pol.W / st->level), pol.n_sms);
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" int rf_track_vertex_normal(const float* depth, int H, int W, const float K[9], float cut_dist, float trunc, int seed,
                                      float sample_range, float* row_sample, float* depth_vertex, float* normal, void* stream) {
    RF_REQUIRE(depth && K && row_sample && depth_vertex && normal, RF_E_NULL, "rf_track_vertex_normal: NULL pointer");
    RF_REQUIRE(H >= 3 && W >= 3 && (long long)H * W < (1ll << 31), RF_E_RANGE, "rf_track_vertex_normal: bad image size %dx%d", W, H);
    RF_REQUIRE(((uintptr_t)depth_vertex & 15) == 0, RF_E_ALIGN, "rf_track_vertex_normal: depth_vertex must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    TrackCam cam; for (int i = 0; i < 9; ++i) cam.k[i] = K[i];
    // the reference ships the seed inside a float32 array (:447-456) and converts it back with int(): same rounding here
    const unsigned long long seed_u = (unsigned long long)(int)(float)seed;
    track_row_sample_kernel<<<(H + 127) / 128, 128, 0, s>>>(H, seed_u, sample_range, row_sample);
    RF_CHECK_LAUNCH("track_row_sample_kernel");
    const int n = H * W;
    track_vertex_kernel<<<(n + 255) / 256, 256, 0, s>>>(depth, row_sample, cam, H, W, cut_dist, trunc, reinterpret_cast<float4*>(depth_vertex));
    RF_CHECK_LAUNCH("track_vertex_kernel");
    track_normal_kernel<<<(n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const float4*>(depth_vertex), H, W, normal);
    RF_CHECK_LAUNCH("track_normal_kernel");
    return 0;
}

extern "C" int64_t rf_track_fitness_scratch_floats(int n_candidates, int H, int W, int level) {
    if (n_candidates <= 0 || level <= 0) return 0;
    return 2ll * n_candidates * fit_chunks(n_candidates, (H / level) * (W / level));
}

extern "C" int rf_track_fitness(const float* tsdf_vol, const int vol_dim[3], const float vol_origin[3], float voxel_size,
                                const float* depth_vertex, const float* normal, int H, int W, const float K[9],
                                const float R[9], const float T[3], const float* candidates, int n_candidates,
                                const float search_size[6], int level, int level_index,
                                float* search_value, float* search_count, float* scratch, void* stream) {
    RF_REQUIRE(tsdf_vol && vol_dim && vol_origin && depth_vertex && normal && K && R && T && candidates && search_size && search_value && search_count && scratch,
               RF_E_NULL, "rf_track_fitness: NULL pointer");
    RF_REQUIRE(n_candidates > 0 && level > 0 && level_index >= 0 && H > 0 && W > 0, RF_E_RANGE, "rf_track_fitness: bad sizes");
    RF_REQUIRE((((uintptr_t)depth_vertex & 15) | ((uintptr_t)scratch & 7)) == 0, RF_E_ALIGN, "rf_track_fitness: depth_vertex 16-byte / scratch 8-byte alignment");
    RF_REQUIRE((long long)vol_dim[0] * vol_dim[1] * vol_dim[2] < (1ll << 31), RF_E_RANGE, "rf_track_fitness: volume too large for the reference's int index");
    FitArgs a;
    a.tsdf = tsdf_vol; a.dx = vol_dim[0]; a.dy = vol_dim[1]; a.dz = vol_dim[2];
    a.ox = (int)vol_origin[0]; a.oy = (int)vol_origin[1]; a.oz = (int)vol_origin[2];      // :163-165: origin truncated to int
    a.voxel = voxel_size;
    a.vertex = reinterpret_cast<const float4*>(depth_vertex); a.normal = normal; a.H = H; a.W = W;
    for (int i = 0; i < 9; ++i) { a.cam.k[i] = K[i]; a.pose.R[i] = R[i]; }
    for (int i = 0; i < 3; ++i) a.pose.T[i] = T[i];
    for (int i = 0; i < 6; ++i) a.pose.ss[i] = search_size[i];
    a.cand = candidates; a.n = n_candidates; a.level = level; a.level_index = level_index; a.st = nullptr;
    a.ph = H / level; a.pw = W / level;                                                    // host :587-588: int(im_h/level)
    const int pixels = a.ph * a.pw;
    a.chunks = fit_chunks(n_candidates, pixels);
    a.pix_per_chunk = std::max(1, (pixels + a.chunks - 1) / a.chunks);
    cudaStream_t s = (cudaStream_t)stream;
    {
        ProfScope ps(RF_PROF_TRACK_FITNESS, s);
        track_fitness_kernel<false><<<dim3((n_candidates + kCand - 1) / kCand, a.chunks), kCand, 0, s>>>(a, scratch);
    }
    RF_CHECK_LAUNCH("track_fitness_kernel");
    track_fold_kernel<<<(n_candidates + 255) / 256, 256, 0, s>>>(scratch, n_candidates, a.chunks, search_value, search_count, nullptr);
    RF_CHECK_LAUNCH("track_fold_kernel");
    return 0;
}

extern "C" int rf_track_cal_transform(const float* search_value, const float* search_count, const float* candidates, int n_candidates,
                                      const float search_size[6], int count_search, float* out9, void* stream) {
    RF_REQUIRE(search_value && search_count && candidates && search_size && out9, RF_E_NULL, "rf_track_cal_transform: NULL pointer");
    RF_REQUIRE(n_candidates >= 1 && count_search >= 0, RF_E_RANGE, "rf_track_cal_transform: bad sizes");
    TrackPose pose; memset(&pose, 0, sizeof(pose));
    for (int i = 0; i < 6; ++i) pose.ss[i] = search_size[i];
    track_cal_transform_kernel<<<1, kCalThreads, 0, (cudaStream_t)stream>>>(search_value, search_count, candidates, n_candidates, pose, count_search, out9, nullptr);
    RF_CHECK_LAUNCH("track_cal_transform_kernel");
    return 0;
}

// ---- the whole search loop on the device (model/ROtracker.py:716-836 random_optimization) -------------------------------------
static int ro_geometry(const int pst_n[20], const int depth_level[20], int H, int W, int& n_max, int& chunks_max) {
    n_max = 0; chunks_max = 1;
    for (int k = 0; k < 20; ++k) {
        if (pst_n[k] <= 0 || depth_level[k] <= 0) return -1;
        n_max = std::max(n_max, pst_n[k]);
        chunks_max = std::max(chunks_max, fit_chunks(pst_n[k], (H / depth_level[k]) * (W / depth_level[k])));
    }
    return 0;
}

extern "C" int64_t rf_track_random_optimization_scratch_floats(const int pst_n[20], const int depth_level[20], int H, int W) {
    int n_max, chunks_max;
    if (!pst_n || !depth_level || ro_geometry(pst_n, depth_level, H, W, n_max, chunks_max)) return 0;
    return 2ll * n_max * chunks_max;
}

extern "C" int rf_track_random_optimization(const float* tsdf_vol, const int vol_dim[3], const float vol_origin[3], float voxel_size,
                                            const float* depth_vertex, const float* normal, int H, int W, const float K[9],
                                            const float* pst, const int pst_offset[20], const int pst_n[20], const int depth_level[20],
                                            int iters, int count_search, float scaling_coefficient, int fix_level_index,
                                            int iterative_scale, float beta, float* state, float* search_value, float* search_count,
                                            float* scratch, void* stream) {
    RF_REQUIRE(tsdf_vol && vol_dim && vol_origin && depth_vertex && normal && K && pst && pst_offset && pst_n && depth_level && state &&
               search_value && search_count && scratch, RF_E_NULL, "rf_track_random_optimization: NULL pointer");
    RF_REQUIRE(iters >= 1 && iters <= 31 && count_search >= 0 && H > 0 && W > 0, RF_E_RANGE, "rf_track_random_optimization: bad sizes");
    RF_REQUIRE((((uintptr_t)depth_vertex & 15) | ((uintptr_t)scratch & 7) | ((uintptr_t)state & 15)) == 0, RF_E_ALIGN,
               "rf_track_random_optimization: depth_vertex / state 16-byte, scratch 8-byte alignment");
    RF_REQUIRE((long long)vol_dim[0] * vol_dim[1] * vol_dim[2] < (1ll << 31), RF_E_RANGE, "rf_track_random_optimization: volume too large for the reference's int index");
    int n_max, chunks_max;
    RF_REQUIRE(ro_geometry(pst_n, depth_level, H, W, n_max, chunks_max) == 0, RF_E_RANGE, "rf_track_random_optimization: candidate counts and pyramid levels must be positive");
    FitArgs a; memset(&a, 0, sizeof(a));
    a.tsdf = tsdf_vol; a.dx = vol_dim[0]; a.dy = vol_dim[1]; a.dz = vol_dim[2];
    a.ox = (int)vol_origin[0]; a.oy = (int)vol_origin[1]; a.oz = (int)vol_origin[2];
    a.voxel = voxel_size;
    a.vertex = reinterpret_cast<const float4*>(depth_vertex); a.normal = normal; a.H = H; a.W = W;
    for (int i = 0; i < 9; ++i) a.cam.k[i] = K[i];
    a.cand = pst;
    TrackDevState* st = reinterpret_cast<TrackDevState*>(state);
    a.st = st;
    TrackPolicy pol; memset(&pol, 0, sizeof(pol));
    pol.iters = iters; pol.count_search = count_search; pol.fix_level_index = fix_level_index; pol.iterative_scale = iterative_scale;
    pol.n_sms = num_sms(); pol.H = H; pol.W = W; pol.scaling = scaling_coefficient; pol.beta = beta;
    for (int k = 0; k < 20; ++k) { pol.pst_n[k] = pst_n[k]; pol.pst_off[k] = pst_offset[k]; pol.depth_level[k] = depth_level[k]; }
    TrackPose none; memset(&none, 0, sizeof(none));
    cudaStream_t s = (cudaStream_t)stream;
    for (int i = 0; i < iters; ++i) {                        // 4 launches per iteration, nothing read back in between
        track_fitness_kernel<true><<<dim3((n_max + kCand - 1) / kCand, chunks_max), kCand, 0, s>>>(a, scratch);
        RF_CHECK_LAUNCH("track_fitness_kernel<dev>");
        track_fold_kernel<<<(n_max + 255) / 256, 256, 0, s>>>(scratch, 0, 0, search_value, search_count, st);
        RF_CHECK_LAUNCH("track_fold_kernel<dev>");
        track_cal_transform_kernel<<<1, kCalThreads, 0, s>>>(search_value, search_count, pst, 1, none, count_search, nullptr, st);
        RF_CHECK_LAUNCH("track_cal_transform_kernel<dev>");
        track_policy_kernel<<<1, 32, 0, s>>>(st, pol);
        RF_CHECK_LAUNCH("track_policy_kernel");
    }
    return 0;
}

extern "C" int rf_track_state_init(float* state_host, const float R[9], const float T[3], const float search_size[6],
                                   const float previous_search_size[6], const int pst_offset[20], const int pst_n[20],
                                   const int depth_level[20], int H, int W) {
    RF_REQUIRE(state_host && R && T && search_size && previous_search_size && pst_offset && pst_n && depth_level, RF_E_NULL, "rf_track_state_init: NULL pointer");
    RF_REQUIRE(pst_n[0] > 0 && depth_level[0] > 0 && H > 0 && W > 0, RF_E_RANGE, "rf_track_state_init: bad sizes");
    TrackDevState st; memset(&st, 0, sizeof(st));
    for (int i = 0; i < 9; ++i) st.R[i] = R[i];
    for (int i = 0; i < 3; ++i) st.T[i] = T[i];
    for (int i = 0; i < 6; ++i) { st.ss[i] = search_size[i]; st.prev_ss[i] = previous_search_size[i]; }
    st.count_particle = 0; st.level_index = 5;                                   // model/ROtracker.py:752-755
    st.n = pst_n[0]; st.level = depth_level[0]; st.cand_off = pst_offset[0];
    st.chunks = fit_chunks(st.n, (H / st.level) * (W / st.level));
    memcpy(state_host, &st, sizeof(st));
    return 0;
}
