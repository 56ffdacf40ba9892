// ray_query.cu — Stage 2 of the mapping hot path: mixed-representation ray query, render, losses, backward.
//
// Replaces, fused, what the reference runs as ~40 ATen launches + 3 tiny-cuda-nn launches + 4 sgemms forward and
// about twice that backward:
//   JointEncoding.render_rays      model/scene_rep.py:407-456   -> ray_z_kernel (sampling) + sample_fwd_kernel + composite_fwd_kernel
//   JointEncoding.run_network      model/scene_rep.py:370-402   -> fp64 normalisation inside the sample kernels
//   JointEncoding.query_color_sdf  model/scene_rep.py:314-349   -> sample_fwd_kernel (hash + OneBlob + GBV trilerp + decoder)
//   ColorSDFNet / SDFNet / ColorNet model/decoder.py:6-146      -> decoder inside the sample kernels
//   raw2outputs / sdf2weights      model/scene_rep.py:107-179   -> composite_fwd_kernel
//   mapping() losses               model/scene_rep.py:493-517, model/utils.py:170-256 -> loss partial sums in composite_fwd_kernel
//   autograd backward (R8)                                       -> composite_bwd_kernel + sample_bwd_kernel + ray_grad_kernel
//
// Kernel plan (mlp_precision 0 = fp32 SIMT decoder, the accuracy anchor):
//   * one thread per sample; decoder weights live transposed in shared memory ([in][out]) so that a thread streams its
//     input features through `h[j] += x_k * Wt[k][j]` with broadcast 128-bit shared loads;
//   * no activation ever goes to HBM: per sample the kernels read ray (28 B, L1-resident across the ray's samples),
//     gather 16x8 hash entries (8 B) + 8 GBV voxels (16 B) and write raw (16 B); backward recomputes the forward;
//   * table gradients are scattered with 8-byte vector reductions (RED.ADD.F32x2), decoder weight gradients are
//     tile GEMMs over the 128 samples of a block accumulated in shared memory and flushed once per block.
#include <math.h>
#include <algorithm>
#include "ray_common.cuh"

namespace rf {

// ------------------------------------------------------------------------------------------------------------
// z sampling — model/scene_rep.py:417-441.  One warp per ray.
// tables: [n_range_d] offsets (linspace(-range_d, range_d)), [n_range_d] linspace(near, far, n_range_d),
//         [n_samples_d] linspace(near, far, n_samples_d) — produced by torch.linspace on the host side so that the
//         values are the ones the reference computes on the same machine.
__global__ void ray_z_kernel(RayK k, const float* __restrict__ target_d, const float* __restrict__ u,
                             const float* __restrict__ tables, float* __restrict__ z_vals) {
    __shared__ float s_z[4][kMaxS];
    __shared__ float s_a[4][kMaxS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long r = blockIdx.x * 4ll + warp;
    if (r >= k.n_rays) return;
    const int nA = k.n_range_d, nB = k.n_samples_d, S = k.S;
    const float* tA = tables; const float* tAnf = tables + nA; const float* tB = tables + 2 * nA;
    float d = target_d[r];
    float* za = s_a[warp]; float* zs = s_z[warp];
    for (int i = lane; i < nA; i += 32) za[i] = (d <= 0.f) ? tAnf[i] : __fadd_rn(tA[i], d);      // :422-424
    __syncwarp();
    if (nB > 0) {                                                                                 // :426-428 (cat + sort)
        // both lists are ascending (linspace; linspace + d), so the rank of an element in the other list is a binary search: the merged
        // position is the same as with the linear counts (#{tB < v} for the near-surface samples, #{za <= v} for the uniform ones)
        for (int i = lane; i < nA; i += 32) {
            const float v = za[i];
            int lo = 0, hi = nB;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (tB[mid] < v) lo = mid + 1; else hi = mid; }
            zs[i + lo] = v;
        }
        for (int j = lane; j < nB; j += 32) {
            const float v = tB[j];
            int lo = 0, hi = nA;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (za[mid] <= v) lo = mid + 1; else hi = mid; }
            zs[j + lo] = v;
        }
    } else {
        for (int i = lane; i < nA; i += 32) zs[i] = za[i];
    }
    __syncwarp();
    for (int s = lane; s < S; s += 32) {
        float z = zs[s];
        if (k.perturb) {                                                                          // :437-441
            float lower = (s == 0) ? zs[0] : __fmul_rn(0.5f, __fadd_rn(zs[s], zs[s - 1]));
            float upper = (s == S - 1) ? zs[S - 1] : __fmul_rn(0.5f, __fadd_rn(zs[s + 1], zs[s]));
            z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u[r * S + s]));
        }
        z_vals[r * S + s] = z;
    }
}

// shared-memory layout of the transposed decoder weights
template <int HID>
struct WSmem {
    float* w0t;   // [in1][HID]
    float* w1t;   // [HID][16]
    float* w2t;   // [66][HID]
    float* w3t;   // [HID][4]
    __device__ static int floats(int in1) { return in1 * HID + HID * kOut1 + kIn2 * HID + HID * 4; }
    __device__ void carve(float* base, int in1) { w0t = base; w1t = w0t + in1 * HID; w2t = w1t + HID * kOut1; w3t = w2t + kIn2 * HID; }
    __device__ void load(const Weights& w, int in1) {
        for (int i = threadIdx.x; i < HID * in1; i += blockDim.x) { int j = i / in1, kk = i - j * in1; w0t[kk * HID + j] = w.w_sdf0[i]; }
        for (int i = threadIdx.x; i < kOut1 * HID; i += blockDim.x) { int o = i / HID, j = i - o * HID; w1t[j * kOut1 + o] = w.w_sdf1[i]; }
        for (int i = threadIdx.x; i < HID * kIn2; i += blockDim.x) { int j = i / kIn2, kk = i - j * kIn2; w2t[kk * HID + j] = w.w_col0[i]; }
        for (int i = threadIdx.x; i < HID * 4; i += blockDim.x) { int j = i >> 2, c = i & 3; w3t[i] = (c < 3) ? w.w_col1[c * HID + j] : 0.f; }
    }
};

// ------------------------------------------------------------------------------------------------------------
// Forward: one thread per sample.
// FROM_X: positions given directly in normalised coordinates (point query) instead of rays + z.
// ------------------------------------------------------------------------------------------------------------
template <int HID, bool FROM_X>
__global__ void __launch_bounds__(kTile) sample_fwd_kernel(RayK k, GridDev hg, GridDev gg, const float* __restrict__ hash_params,
                                                           const float* __restrict__ gbv_params, Weights wts,
                                                           const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                           const float* __restrict__ z_vals, const float* __restrict__ xin,
                                                           long long P, int variant, float* __restrict__ raw) {
    extern __shared__ __align__(16) float smem[];
    WSmem<HID> W; W.carve(smem, k.in1); W.load(wts, k.in1);
    __syncthreads();
    const int L = hg.n_levels;
    for (long long p = blockIdx.x * (long long)kTile + threadIdx.x; p < P; p += (long long)gridDim.x * kTile) {
        float x[3];
        if (FROM_X) { x[0] = xin[3 * p]; x[1] = xin[3 * p + 1]; x[2] = xin[3 * p + 2]; }
        else sample_x(k, rays_o, rays_d, p / k.S, z_vals[p], x);
        float h[HID];
#pragma unroll
        for (int j = 0; j < HID; ++j) h[j] = 0.f;
        for (int l = 0; l < L; ++l) {                                       // model/scene_rep.py:325
            float2 f = hash_level_feat(hg, hash_params, l, x);
            axpy_row<HID>(h, f.x, W.w0t + (2 * l) * HID);
            axpy_row<HID>(h, f.y, W.w0t + (2 * l + 1) * HID);
        }
        const int ob0 = k.n_hash_out;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {                                       // :327
            float ob[kNB];
            oneblob_coord<kNB>(x[c], ob);
#pragma unroll
            for (int b = 0; b < kNB; ++b) axpy_row<HID>(h, ob[b], W.w0t + (ob0 + c * kNB + b) * HID);
        }
        float4 g = gbv_feat(gg, gbv_params, x);                             // :329
        float t_add, cin, d0, d1;
        tsdf_terms(k, variant, g.x, t_add, cin, d0, d1);
        axpy_row<HID>(h, cin, W.w0t + (ob0 + kBlob) * HID);
        float o16[kOut1];
#pragma unroll
        for (int i = 0; i < kOut1; ++i) o16[i] = 0.f;
#pragma unroll
        for (int j = 0; j < HID; ++j) {                                     // model/decoder.py:105-110 (ReLU, Linear)
            float a = fmaxf(h[j], 0.f);
            const float4* wr = reinterpret_cast<const float4*>(W.w1t + j * kOut1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v = wr[q];
                o16[4 * q] = fmaf(a, v.x, o16[4 * q]); o16[4 * q + 1] = fmaf(a, v.y, o16[4 * q + 1]);
                o16[4 * q + 2] = fmaf(a, v.z, o16[4 * q + 2]); o16[4 * q + 3] = fmaf(a, v.w, o16[4 * q + 3]);
            }
        }
        // colour net: input = [oneblob48, geo15, gbv_rgb3]  (model/decoder.py:141)
#pragma unroll
        for (int j = 0; j < HID; ++j) h[j] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            float ob[kNB];
            oneblob_coord<kNB>(x[c], ob);
#pragma unroll
            for (int b = 0; b < kNB; ++b) axpy_row<HID>(h, ob[b], W.w2t + (c * kNB + b) * HID);
        }
#pragma unroll
        for (int i = 0; i < kGeo; ++i) axpy_row<HID>(h, o16[1 + i], W.w2t + (kBlob + i) * HID);
        axpy_row<HID>(h, g.y, W.w2t + (kBlob + kGeo) * HID);
        axpy_row<HID>(h, g.z, W.w2t + (kBlob + kGeo + 1) * HID);
        axpy_row<HID>(h, g.w, W.w2t + (kBlob + kGeo + 2) * HID);
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            float a = fmaxf(h[j], 0.f);
            float4 v = *reinterpret_cast<const float4*>(W.w3t + 4 * j);
            r0 = fmaf(a, v.x, r0); r1 = fmaf(a, v.y, r1); r2 = fmaf(a, v.z, r2);
        }
        // :344-345 residual add
        reinterpret_cast<float4*>(raw)[p] = make_float4(r0 + g.y, r1 + g.z, r2 + g.w, o16[0] + t_add);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Composite (sdf2weights + raw2outputs, model/scene_rep.py:107-127,156-179) and the loss partial sums
// (model/scene_rep.py:493-517, model/utils.py:170-256).  One warp per ray.
// ------------------------------------------------------------------------------------------------------------
// sigmoid(a) and the unnormalised rendering weight sigmoid(a) sigmoid(-a) (model/scene_rep.py:116) = t / (1 + t)^2 with
// t = exp(-|a|): one exponential (ex2.approx) and one reciprocal (rcp.approx; 1 + t lies in [1, 2]) instead of two IEEE
// divisions and two expf per sample — the composite kernels were bound by exactly that instruction stream.  Relative error of
// both results <= 8e-7 + 6e-8 |a| (tests/test_ray_oracle.py models it; the argument of ex2.approx is rounded once; the weights that far from the surface are
// e^-|a| small); parity bar of the rendered outputs: 2e-4, observed agreement with the oracle unchanged (profiles/r2l_parity_stats.jsonl).
__device__ __forceinline__ void sdf_weight(float a, float& sg, float& e) {
    const float t = __expf(-fabsf(a));
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(1.0f + t));
    const float ti = t * inv;
    sg = (a >= 0.f) ? inv : ti;
    e = ti * inv;
}

struct RayState { int first; float zthr; float denom; float inv_denom; };

// A ray's samples in registers: lane l holds samples l, l + 32, ... (S <= kMaxS = 128: at most four), with the
// unnormalised weight e = sigmoid(s / trunc) sigmoid(-s / trunc) (model/scene_rep.py:116) computed once.
// The kernels are instantiated per NPL = ceil(S / 32) (1..4), so that a 59-sample ray keeps two samples per lane in registers, not four.
template <int NPL> struct RaySamples { float4 v[NPL]; float z[NPL]; float sg[NPL]; float e[NPL]; };

template <int NPL>
__device__ __forceinline__ void load_ray(const RayK& k, float inv_trunc, const float4* __restrict__ raw_r, const float* __restrict__ z_r, int lane, RaySamples<NPL>& rs) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int s = lane + 32 * i;
        if (s < k.S) {
            rs.v[i] = raw_r[s]; rs.z[i] = z_r[s];
            sdf_weight(rs.v[i].w * inv_trunc, rs.sg[i], rs.e[i]);
        } else {
            rs.v[i] = make_float4(0.f, 0.f, 0.f, 0.f); rs.z[i] = 0.f; rs.sg[i] = 0.f; rs.e[i] = 0.f;
        }
    }
}

// first sign change (argmax of the 0/1 mask => 0 if none), truncation threshold, normaliser (:119-127)
template <int NPL>
__device__ __forceinline__ RayState ray_state(const RayK& k, const float4* __restrict__ raw_r, const float* __restrict__ z_r, int lane,
                                              const RaySamples<NPL>& rs) {
    const int S = k.S;
    int first = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int s = lane + 32 * i;
        if (s < S - 1 && first == 0x7fffffff && raw_r[s + 1].w * rs.v[i].w < 0.f) first = s;
    }
    first = __reduce_min_sync(0xffffffffu, first);
    if (first == 0x7fffffff) first = 0;
    RayState st; st.first = first;
    st.zthr = __fadd_rn(z_r[first], k.sc_trunc);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int s = lane + 32 * i;
        if (s < S) sum += (rs.z[i] < st.zthr) ? rs.e[i] : 0.f;
    }
    st.denom = warp_sum(sum) + 1e-8f;
    st.inv_denom = __frcp_rn(st.denom);              // w = e / denom as one reciprocal per ray (<= 1 ulp from the quotient)
    return st;
}

// Warp per ray, four rays per block; the seven loss partial sums (double) are added per block.  (A persistent variant with
// one set of atomics per block of many rays was slower: 91 registers, 1.02 vs 0.91 ms — the atomics are not the bound.)
template <int NPL>
__global__ void __launch_bounds__(128, NPL == 4 ? 10 : 12) composite_fwd_kernel(RayK k, const float* __restrict__ raw, const float* __restrict__ z_vals,
                                                            const float* __restrict__ target_d, const float* __restrict__ target_rgb,
                                                            float* __restrict__ rgb_map, float* __restrict__ depth_map, double* __restrict__ partials) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    const int S = k.S;
    const long long r = blockIdx.x * 4ll + warp;
    if (r < k.n_rays) {
        const float4* raw_r = reinterpret_cast<const float4*>(raw) + r * S;
        const float* z_r = z_vals + r * S;
        RaySamples<NPL> rs; load_ray(k, __frcp_rn(k.trunc), raw_r, z_r, lane, rs);
        RayState st = ray_state(k, raw_r, z_r, lane, rs);
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, dm = 0.f;
        float d = partials ? target_d[r] : 0.f;
        bool valid = (d > 0.f) && (d < k.depth_trunc);
        float fs = 0.f, sd = 0.f; int nf = 0, ns = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            if (lane + 32 * i >= S) continue;
            const float4 v = rs.v[i]; const float z = rs.z[i];
            float w = (z < st.zthr) ? rs.e[i] : 0.f;
            w = w * st.inv_denom;
            c0 = fmaf(w, v.x, c0); c1 = fmaf(w, v.y, c1); c2 = fmaf(w, v.z, c2); dm = fmaf(w, z, dm);
            if (partials) {
                bool front = z < __fsub_rn(d, k.sc_trunc), back = z > __fadd_rn(d, k.sc_trunc);
                bool sm = !front && !back && (d > 0.f);
                nf += front ? 1 : 0; ns += sm ? 1 : 0;
                if (valid && front) { float e = v.w - 1.0f; fs = fmaf(e, e, fs); }
                if (valid && sm) { float e = __fadd_rn(z, __fmul_rn(v.w, k.sc_trunc)) - d; sd = fmaf(e, e, sd); }
            }
        }
        c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); dm = warp_sum(dm);
        if (lane == 0) { rgb_map[3 * r] = c0; rgb_map[3 * r + 1] = c1; rgb_map[3 * r + 2] = c2; depth_map[r] = dm; }
        if (partials) {
            fs = warp_sum(fs); sd = warp_sum(sd); nf = __reduce_add_sync(0xffffffffu, nf); ns = __reduce_add_sync(0xffffffffu, ns);
            if (lane == 0) {
                float wgt = (k.rgb_all_ones || valid) ? 1.f : 0.f;
                float e0 = c0 * wgt - target_rgb[3 * r] * wgt, e1 = c1 * wgt - target_rgb[3 * r + 1] * wgt, e2 = c2 * wgt - target_rgb[3 * r + 2] * wgt;
                acc[0] = (double)e0 * e0 + (double)e1 * e1 + (double)e2 * e2;
                if (valid) { float e = dm - d; acc[1] = (double)e * e; acc[2] = 1.0; }
                acc[3] = fs; acc[4] = sd; acc[5] = nf; acc[6] = ns;
            }
        }
    }
    if (partials) {
        __shared__ double s_acc[4][7];
        if (lane == 0) for (int i = 0; i < 7; ++i) s_acc[warp][i] = acc[i];
        __syncthreads();
        if (threadIdx.x < 7) {
            double v = s_acc[0][threadIdx.x] + s_acc[1][threadIdx.x] + s_acc[2][threadIdx.x] + s_acc[3][threadIdx.x];
            if (v != 0.0) atomicAdd(partials + threadIdx.x, v);
        }
    }
}

// Backward of composite + losses: writes the total gradient w.r.t. raw [N,S,4] into d_raw_out.
// Upstream: d_rgb_map [N,3], d_depth_map [N], d_raw [N,S,4] (each may be NULL) and loss_grads (device float[4]:
// d/d rgb_loss, depth_loss, sdf_loss, fs_loss; may be NULL) with the forward's `partials`.
template <int NPL>
__global__ void __launch_bounds__(128, NPL == 4 ? 7 : 8) composite_bwd_kernel(RayK k, const float* __restrict__ raw, const float* __restrict__ z_vals,
                                                            const float* __restrict__ rgb_map, const float* __restrict__ depth_map,
                                                            const float* __restrict__ target_d, const float* __restrict__ target_rgb,
                                                            const float* __restrict__ d_rgb_map, const float* __restrict__ d_depth_map,
                                                            const float* __restrict__ d_raw, const float* __restrict__ loss_grads,
                                                            const double* __restrict__ partials, float* __restrict__ d_raw_out,
                                                            int* __restrict__ n_live) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long r = blockIdx.x * 4ll + warp;
    if (r >= k.n_rays) return;
    const int S = k.S;
    const float4* raw_r = reinterpret_cast<const float4*>(raw) + r * S;
    const float* z_r = z_vals + r * S;
    const float inv_trunc = __frcp_rn(k.trunc);
    RaySamples<NPL> rs; load_ray(k, inv_trunc, raw_r, z_r, lane, rs);
    RayState st = ray_state(k, raw_r, z_r, lane, rs);
    float G0 = 0.f, G1 = 0.f, G2 = 0.f, GD = 0.f;
    if (d_rgb_map) { G0 = d_rgb_map[3 * r]; G1 = d_rgb_map[3 * r + 1]; G2 = d_rgb_map[3 * r + 2]; }
    if (d_depth_map) GD = d_depth_map[r];
    float d = 0.f; bool valid = false; float cfs = 0.f, csd = 0.f;
    if (loss_grads) {
        d = target_d[r];
        valid = (d > 0.f) && (d < k.depth_trunc);
        const double NT = (double)k.n_total;
        float wgt = (k.rgb_all_ones || valid) ? 1.f : 0.f;
        float gl = loss_grads[0] * (float)(2.0 / (3.0 * NT));                          // mse over N*3 elements
        G0 += gl * (rgb_map[3 * r] * wgt - target_rgb[3 * r] * wgt) * wgt;
        G1 += gl * (rgb_map[3 * r + 1] * wgt - target_rgb[3 * r + 1] * wgt) * wgt;
        G2 += gl * (rgb_map[3 * r + 2] * wgt - target_rgb[3 * r + 2] * wgt) * wgt;
        double nv = partials[2];
        if (valid && nv > 0) GD += loss_grads[1] * (float)(2.0 / nv) * (depth_map[r] - d);
        double nf = partials[5], ns = partials[6], nn = nf + ns;                        // model/utils.py:190-196
        double fs_w = 1.0 - nf / nn, sdf_w = 1.0 - ns / nn;
        cfs = loss_grads[3] * (float)(fs_w * 2.0 / (NT * S));
        csd = loss_grads[2] * (float)(sdf_w * 2.0 / (NT * S));
    }
    // pass 1: dot = sum_j dw_j * w_j
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        if (lane + 32 * i >= S) continue;
        const float4 v = rs.v[i]; const float z = rs.z[i];
        float w = ((z < st.zthr) ? rs.e[i] : 0.f) * st.inv_denom;
        float dw = G0 * v.x + G1 * v.y + G2 * v.z + GD * z;
        dot = fmaf(dw, w, dot);
    }
    dot = warp_sum(dot);
    float4* out_r = reinterpret_cast<float4*>(d_raw_out) + r * S;
    int last_nz = 0;
    const float4* up_r = d_raw ? reinterpret_cast<const float4*>(d_raw) + r * S : nullptr;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int s = lane + 32 * i;
        if (s >= S) continue;
        const float4 v = rs.v[i]; const float z = rs.z[i];
        float sg = rs.sg[i];
        float e = (z < st.zthr) ? rs.e[i] : 0.f;
        float w = e * st.inv_denom;
        float dw = G0 * v.x + G1 * v.y + G2 * v.z + GD * z;
        float de = (dw - dot) * st.inv_denom;
        float4 o;
        o.x = w * G0; o.y = w * G1; o.z = w * G2;
        o.w = de * e * (1.0f - 2.0f * sg) * inv_trunc;
        if (loss_grads) {
            bool front = z < __fsub_rn(d, k.sc_trunc), back = z > __fadd_rn(d, k.sc_trunc);
            bool sm = !front && !back && (d > 0.f);
            if (valid && front) o.w += cfs * (v.w - 1.0f);
            if (valid && sm) o.w += csd * (__fadd_rn(z, __fmul_rn(v.w, k.sc_trunc)) - d) * k.sc_trunc;
        }
        if (up_r) { float4 uu = up_r[s]; o.x += uu.x; o.y += uu.y; o.z += uu.z; o.w += uu.w; }
        out_r[s] = o;
        if (o.x != 0.f || o.y != 0.f || o.z != 0.f || o.w != 0.f) last_nz = s + 1;
    }
    // n_live: one past the last sample with a non-zero gradient.  Samples behind the truncation band have zero rendering
    // weight (scene_rep.py:124) and no loss term (utils.py:170-198), so the tail of most rays is exactly zero and the
    // decoder backward / table scatter skip it.
    if (n_live) {
        last_nz = __reduce_max_sync(0xffffffffu, last_nz);
        if (lane == 0) n_live[r] = last_nz;
    }
}

// one instantiation per number of samples a lane holds (S <= kMaxS = 128)
static void launch_composite_fwd(const RayK& k, const float* raw, const float* z_vals, const float* target_d, const float* target_rgb,
                                 float* rgb_map, float* depth_map, double* partials, cudaStream_t s) {
    const unsigned grid = (unsigned)((k.n_rays + 3) / 4);
    switch ((k.S + 31) / 32) {
        case 1: composite_fwd_kernel<1><<<grid, 128, 0, s>>>(k, raw, z_vals, target_d, target_rgb, rgb_map, depth_map, partials); break;
        case 2: composite_fwd_kernel<2><<<grid, 128, 0, s>>>(k, raw, z_vals, target_d, target_rgb, rgb_map, depth_map, partials); break;
        case 3: composite_fwd_kernel<3><<<grid, 128, 0, s>>>(k, raw, z_vals, target_d, target_rgb, rgb_map, depth_map, partials); break;
        default: composite_fwd_kernel<4><<<grid, 128, 0, s>>>(k, raw, z_vals, target_d, target_rgb, rgb_map, depth_map, partials); break;
    }
}
static void launch_composite_bwd(const RayK& k, const float* raw, const float* z_vals, const float* rgb_map, const float* depth_map,
                                 const float* target_d, const float* target_rgb, const float* d_rgb_map, const float* d_depth_map,
                                 const float* d_raw, const float* loss_grads, const double* partials, float* d_raw_out, int* n_live, cudaStream_t s) {
    const unsigned grid = (unsigned)((k.n_rays + 3) / 4);
#define RF_CBWD(N) composite_bwd_kernel<N><<<grid, 128, 0, s>>>(k, raw, z_vals, rgb_map, depth_map, target_d, target_rgb, d_rgb_map, d_depth_map, \
                                                                d_raw, loss_grads, partials, d_raw_out, n_live)
    switch ((k.S + 31) / 32) {
        case 1: RF_CBWD(1); break;
        case 2: RF_CBWD(2); break;
        case 3: RF_CBWD(3); break;
        default: RF_CBWD(4); break;
    }
#undef RF_CBWD
}

// ------------------------------------------------------------------------------------------------------------
// Sample backward: recompute the forward of a 128-sample tile, back-propagate through the decoder, scatter the
// hash-table gradient, accumulate decoder weight gradients (tile GEMMs in shared memory), optional d/d position.
// ------------------------------------------------------------------------------------------------------------
// dW[n][k0..k0+KPT) += sum_m A[m][n] * B[m][k]   for the 128 rows of the tile; 128 threads.
template <int NA, int KPT>
__device__ __forceinline__ void tile_gemm_tn(float* __restrict__ dW, int ldw, const float* __restrict__ A, int lda,
                                             const float* __restrict__ B, int ldb, int K) {
    const int n = threadIdx.x % NA, k0 = (threadIdx.x / NA) * KPT;
    if (k0 >= K) return;
    float acc[KPT];
#pragma unroll
    for (int i = 0; i < KPT; ++i) acc[i] = 0.f;
    for (int m = 0; m < kTile; ++m) {
        float a = A[m * lda + n];
        const float* b = B + m * ldb + k0;
#pragma unroll
        for (int i = 0; i < KPT; ++i) if (k0 + i < K) acc[i] = fmaf(a, b[i], acc[i]);
    }
#pragma unroll
    for (int i = 0; i < KPT; ++i) if (k0 + i < K) dW[n * ldw + k0 + i] += acc[i];
}


template <int HID, bool BA>
__global__ void __launch_bounds__(kTile) sample_bwd_kernel(RayK k, GridDev hg, GridDev gg, const float* __restrict__ hash_params,
                                                           const float* __restrict__ gbv_params, Weights wts,
                                                           const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                           const float* __restrict__ z_vals, long long P,
                                                           const float* __restrict__ d_raw_tot, Grads gr, float* __restrict__ d_pts) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDX1 = 84;                 // >= 81, multiple of 4
    constexpr int LDX2 = 20;                 // geo15 + rgb3 (+2 pad)
    constexpr int GROUPS = kTile / HID;      // k-groups of the weight-gradient GEMMs
    const int in1 = k.in1;
    WSmem<HID> W; W.carve(smem, in1);
    float* dw0 = smem + WSmem<HID>::floats(in1);          // [HID][in1]   (nn.Linear layout [out][in])
    float* dw1 = dw0 + HID * in1;                          // [16][HID]
    float* dw2 = dw1 + kOut1 * HID;                        // [HID][66]
    float* dw3 = dw2 + HID * kIn2;                         // [3][HID]
    float* X1  = dw3 + 4 * HID;                            // [128][84]   sdf-net input row
    float* H1r = X1 + kTile * LDX1;                        // [128][HID]  relu(hidden sdf)
    float* Rg  = H1r + kTile * HID;                        // phase region
    float* X2s = Rg;                                       // [128][20]   colour-net inputs 48..65
    float* DH2 = X2s + kTile * LDX2;                       // [128][HID]
    float* DO_ = Rg;                                       // [128][16]   (sdf phase overlays the colour phase)
    float* DH1 = DO_ + kTile * kOut1;                      // [128][HID]
    W.load(wts, in1);
    for (int i = threadIdx.x; i < HID * in1 + kOut1 * HID + HID * kIn2 + 4 * HID; i += kTile) dw0[i] = 0.f;
    __syncthreads();
    const int L = hg.n_levels, ob0 = k.n_hash_out;
    const int m = threadIdx.x;
    float* x1 = X1 + m * LDX1;

    for (long long base = blockIdx.x * (long long)kTile; base < P; base += (long long)gridDim.x * kTile) {
        const long long p = base + m;
        const bool live = p < P;
        float x[3] = {0.f, 0.f, 0.f};
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        float t_add = 0.f, cin = 0.f, dg_add = 0.f, dg_cin = 0.f;
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        float h[HID];
#pragma unroll
        for (int j = 0; j < HID; ++j) h[j] = 0.f;
        if (live) {
            sample_x(k, rays_o, rays_d, p / k.S, z_vals[p], x);
            dr = reinterpret_cast<const float4*>(d_raw_tot)[p];
            for (int l = 0; l < L; ++l) {
                float2 f = hash_level_feat(hg, hash_params, l, x);
                x1[2 * l] = f.x; x1[2 * l + 1] = f.y;
                axpy_row<HID>(h, f.x, W.w0t + (2 * l) * HID);
                axpy_row<HID>(h, f.y, W.w0t + (2 * l + 1) * HID);
            }
#pragma unroll 1
            for (int c = 0; c < 3; ++c) {
                float ob[kNB];
                oneblob_coord<kNB>(x[c], ob);
#pragma unroll
                for (int b = 0; b < kNB; ++b) { x1[ob0 + c * kNB + b] = ob[b]; axpy_row<HID>(h, ob[b], W.w0t + (ob0 + c * kNB + b) * HID); }
            }
            g = gbv_feat(gg, gbv_params, x);
            tsdf_terms(k, 0, g.x, t_add, cin, dg_add, dg_cin);
            x1[ob0 + kBlob] = cin;
            axpy_row<HID>(h, cin, W.w0t + (ob0 + kBlob) * HID);
        } else {
            for (int i = 0; i < in1; ++i) x1[i] = 0.f;
        }
        float o16[kOut1];
#pragma unroll
        for (int i = 0; i < kOut1; ++i) o16[i] = 0.f;
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            float a = fmaxf(h[j], 0.f);
            H1r[m * HID + j] = a;
            const float4* wr = reinterpret_cast<const float4*>(W.w1t + j * kOut1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v = wr[q];
                o16[4 * q] = fmaf(a, v.x, o16[4 * q]); o16[4 * q + 1] = fmaf(a, v.y, o16[4 * q + 1]);
                o16[4 * q + 2] = fmaf(a, v.z, o16[4 * q + 2]); o16[4 * q + 3] = fmaf(a, v.w, o16[4 * q + 3]);
            }
        }
        // colour-net forward (hidden pre-activation in h)
        float* x2 = X2s + m * LDX2;
#pragma unroll
        for (int i = 0; i < kGeo; ++i) x2[i] = o16[1 + i];
        x2[kGeo] = g.y; x2[kGeo + 1] = g.z; x2[kGeo + 2] = g.w; x2[kGeo + 3] = 0.f; x2[kGeo + 4] = 0.f;
#pragma unroll
        for (int j = 0; j < HID; ++j) h[j] = 0.f;
        for (int b = 0; b < kBlob; ++b) axpy_row<HID>(h, x1[ob0 + b], W.w2t + b * HID);
#pragma unroll
        for (int i = 0; i < kGeo + 3; ++i) axpy_row<HID>(h, x2[i], W.w2t + (kBlob + i) * HID);
        // colour-net backward: dh2 = relu'(h2) * W3^T d_rgb ; dW3 via warp reduction
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            float4 v = *reinterpret_cast<const float4*>(W.w3t + 4 * j);
            float a = fmaxf(h[j], 0.f);
            float s0 = warp_sum(dr.x * a), s1 = warp_sum(dr.y * a), s2 = warp_sum(dr.z * a);
            if ((threadIdx.x & 31) == 0) { atomicAdd(dw3 + j, s0); atomicAdd(dw3 + HID + j, s1); atomicAdd(dw3 + 2 * HID + j, s2); }
            float dh = (h[j] > 0.f) ? fmaf(v.x, dr.x, fmaf(v.y, dr.y, v.z * dr.z)) : 0.f;
            h[j] = dh;
            DH2[m * HID + j] = dh;
        }
        // d colour-net inputs: geo (always), OneBlob + gbv rgb (BA only)
        float dgeo[kGeo];
#pragma unroll
        for (int i = 0; i < kGeo; ++i) dgeo[i] = dot_row<HID>(h, W.w2t + (kBlob + i) * HID);
        float dx[3] = {0.f, 0.f, 0.f};
        float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);          // gradient w.r.t. the GBV features (tsdf, r, g, b)
        float dob[BA ? kBlob : 1];
        if constexpr (BA) {
            for (int b = 0; b < kBlob; ++b) dob[b] = dot_row<HID>(h, W.w2t + b * HID);
            dg.y = dot_row<HID>(h, W.w2t + (kBlob + kGeo) * HID) + dr.x;          // + residual add (:344)
            dg.z = dot_row<HID>(h, W.w2t + (kBlob + kGeo + 1) * HID) + dr.y;
            dg.w = dot_row<HID>(h, W.w2t + (kBlob + kGeo + 2) * HID) + dr.z;
        }
        __syncthreads();
        // dW2[j][k] += DH2^T [X1(oneblob) | X2s]
        tile_gemm_tn<HID, (kBlob + GROUPS - 1) / GROUPS>(dw2, kIn2, DH2, HID, X1 + ob0, LDX1, kBlob);
        tile_gemm_tn<HID, (kGeo + 3 + GROUPS - 1) / GROUPS>(dw2 + kBlob, kIn2, DH2, HID, X2s, LDX2, kGeo + 3);
        __syncthreads();
        // sdf-net backward
        float do16[kOut1];
        do16[0] = dr.w;
#pragma unroll
        for (int i = 0; i < kGeo; ++i) do16[1 + i] = dgeo[i];
#pragma unroll
        for (int i = 0; i < kOut1; ++i) DO_[m * kOut1 + i] = do16[i];
#pragma unroll
        for (int j = 0; j < HID; ++j) {
            const float4* wr = reinterpret_cast<const float4*>(W.w1t + j * kOut1);
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v = wr[q];
                acc = fmaf(v.x, do16[4 * q], acc); acc = fmaf(v.y, do16[4 * q + 1], acc);
                acc = fmaf(v.z, do16[4 * q + 2], acc); acc = fmaf(v.w, do16[4 * q + 3], acc);
            }
            float dh = (H1r[m * HID + j] > 0.f) ? acc : 0.f;
            h[j] = dh;
            DH1[m * HID + j] = dh;
        }
        if (live) {
            // hash-table gradient scatter (Appendix B5) and, in BA mode, input gradients (B6)
            for (int l = 0; l < L; ++l) {
                float d0 = dot_row<HID>(h, W.w0t + (2 * l) * HID), d1 = dot_row<HID>(h, W.w0t + (2 * l + 1) * HID);
                unsigned cx, cy, cz; float fx, fy, fz;
                pos_fract(x[0], hg.scale[l], cx, fx); pos_fract(x[1], hg.scale[l], cy, fy); pos_fract(x[2], hg.scale[l], cz, fz);
                float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    unsigned idx = grid_index(hg.is_hash, hg.size[l], hg.res[l], cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1));
                    size_t e = (size_t)hg.offset[l] + idx;
                    float w = corner_weight(c, fx, fy, fz);
                    if (gr.g_hash) atomicAdd(reinterpret_cast<float2*>(gr.g_hash) + e, make_float2(w * d0, w * d1));
                    if (BA) {
                        float2 pv = __ldg(reinterpret_cast<const float2*>(hash_params) + e);
                        float v = fmaf(d0, pv.x, d1 * pv.y);
                        float wx = (c & 1) ? fx : 1.f - fx, wy = (c & 2) ? fy : 1.f - fy, wz = (c & 4) ? fz : 1.f - fz;
                        gx += ((c & 1) ? v : -v) * wy * wz; gy += ((c & 2) ? v : -v) * wx * wz; gz += ((c & 4) ? v : -v) * wx * wy;
                    }
                }
                if (BA) { float s = hg.scale[l]; dx[0] = fmaf(gx, s, dx[0]); dx[1] = fmaf(gy, s, dx[1]); dx[2] = fmaf(gz, s, dx[2]); }
            }
            if constexpr (BA) {
                // OneBlob inputs of both nets
                for (int b = 0; b < kBlob; ++b) dob[b] += dot_row<HID>(h, W.w0t + (ob0 + b) * HID);
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {
                    float gb[kNB];
                    oneblob_coord_grad<kNB>(x[c], gb);
                    float acc = 0.f;
#pragma unroll
                    for (int b = 0; b < kNB; ++b) acc = fmaf(gb[b], dob[c * kNB + b], acc);
                    dx[c] += acc;
                }
                // tsdf channel: decoder input (cin) and the residual add (t_add)
                float dcin = dot_row<HID>(h, W.w0t + (ob0 + kBlob) * HID);
                dg.x = dcin * dg_cin + dr.w * dg_add;
                // GBV trilinear input gradient
                unsigned cx, cy, cz; float fx, fy, fz;
                pos_fract(x[0], gg.scale[0], cx, fx); pos_fract(x[1], gg.scale[0], cy, fy); pos_fract(x[2], gg.scale[0], cz, fz);
                float gx = 0.f, gy = 0.f, gz = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    unsigned idx = grid_index(false, gg.size[0], gg.res[0], cx + (c & 1), cy + ((c >> 1) & 1), cz + ((c >> 2) & 1));
                    float4 pv = __ldg(reinterpret_cast<const float4*>(gbv_params) + idx);
                    float v = dg.x * pv.x + dg.y * pv.y + dg.z * pv.z + dg.w * pv.w;
                    float wx = (c & 1) ? fx : 1.f - fx, wy = (c & 2) ? fy : 1.f - fy, wz = (c & 4) ? fz : 1.f - fz;
                    gx += ((c & 1) ? v : -v) * wy * wz; gy += ((c & 2) ? v : -v) * wx * wz; gz += ((c & 4) ? v : -v) * wx * wy;
                }
                float s = gg.scale[0];
                dx[0] = fmaf(gx, s, dx[0]); dx[1] = fmaf(gy, s, dx[1]); dx[2] = fmaf(gz, s, dx[2]);
                // through the float64 normalisation (:388)
                d_pts[3 * p]     = (float)((double)dx[0] / k.bl[0]);
                d_pts[3 * p + 1] = (float)((double)dx[1] / k.bl[1]);
                d_pts[3 * p + 2] = (float)((double)dx[2] / k.bl[2]);
            }
        }
        __syncthreads();
        // dW1[i][j] += DO^T H1r ; dW0[j][k] += DH1^T X1
        tile_gemm_tn<kOut1, HID / (kTile / kOut1)>(dw1, HID, DO_, kOut1, H1r, HID, HID);
        tile_gemm_tn<HID, (81 + GROUPS - 1) / GROUPS>(dw0, in1, DH1, HID, X1, LDX1, in1);
        __syncthreads();
    }
    // flush decoder weight gradients
    if (gr.g_w_sdf0) for (int i = threadIdx.x; i < HID * in1; i += kTile) atomicAdd(gr.g_w_sdf0 + i, dw0[i]);
    if (gr.g_w_sdf1) for (int i = threadIdx.x; i < kOut1 * HID; i += kTile) atomicAdd(gr.g_w_sdf1 + i, dw1[i]);
    if (gr.g_w_col0) for (int i = threadIdx.x; i < HID * kIn2; i += kTile) atomicAdd(gr.g_w_col0 + i, dw2[i]);
    if (gr.g_w_col1) for (int i = threadIdx.x; i < 3 * HID; i += kTile) atomicAdd(gr.g_w_col1 + i, dw3[i]);
}

// dL/d rays_o = sum_s dL/d pts ; dL/d rays_d = sum_s z_s * dL/d pts   (pts = o + d*z, model/scene_rep.py:443)
__global__ void ray_grad_kernel(RayK k, const float* __restrict__ d_pts, const float* __restrict__ z_vals,
                                float* __restrict__ g_o, float* __restrict__ g_d) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long r = blockIdx.x * 4ll + warp;
    if (r >= k.n_rays) return;
    float o0 = 0, o1 = 0, o2 = 0, d0 = 0, d1 = 0, d2 = 0;
    for (int s = lane; s < k.S; s += 32) {
        long long p = r * k.S + s;
        float z = z_vals[p], a = d_pts[3 * p], b = d_pts[3 * p + 1], c = d_pts[3 * p + 2];
        o0 += a; o1 += b; o2 += c; d0 = fmaf(z, a, d0); d1 = fmaf(z, b, d1); d2 = fmaf(z, c, d2);
    }
    o0 = warp_sum(o0); o1 = warp_sum(o1); o2 = warp_sum(o2); d0 = warp_sum(d0); d1 = warp_sum(d1); d2 = warp_sum(d2);
    if (lane == 0) {
        if (g_o) { g_o[3 * r] = o0; g_o[3 * r + 1] = o1; g_o[3 * r + 2] = o2; }
        if (g_d) { g_d[3 * r] = d0; g_d[3 * r + 1] = d1; g_d[3 * r + 2] = d2; }
    }
}

// losses[4] = (rgb, depth, sdf, fs) from the seven partial sums (float64 arithmetic, one thread): mse over N*3 (:501),
// mean over the valid rays (:504-507), the two SDF losses with the front / band balance weights of get_masks
// (model/utils.py:190-196, :242-245)
__global__ void loss_finalize_kernel(const double* __restrict__ p, double n_total, double n_samples, float* __restrict__ losses) {
    const double nn = p[5] + p[6];
    losses[0] = (float)(p[0] / (3.0 * n_total));
    losses[1] = (float)(p[1] / p[2]);
    losses[2] = (float)(p[4] / (n_total * n_samples) * (1.0 - p[6] / nn));
    losses[3] = (float)(p[3] / (n_total * n_samples) * (1.0 - p[5] / nn));
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int make_rayk(RayK& k, const rf_ray_cfg* c, const rf_grid_desc* hash, const rf_grid_desc* gbv, long long n_rays, const char* who) {
    RF_REQUIRE(c && hash && gbv, RF_E_NULL, "%s: NULL config/descriptor", who);
    RF_REQUIRE(hash->n_features == 2 && hash->n_levels >= 1 && hash->n_levels <= RF_MAX_LEVELS, RF_E_UNSUPPORTED, "%s: hash grid must have F=2", who);
    RF_REQUIRE(gbv->n_features == 4 && gbv->n_levels == 1 && !gbv->is_hash, RF_E_UNSUPPORTED, "%s: GBV must be a 1-level dense grid with F=4", who);
    RF_REQUIRE(c->hidden == 32 || c->hidden == 64, RF_E_UNSUPPORTED, "%s: hidden width %d (32 and 64 are built)", who, c->hidden);
    RF_REQUIRE(c->n_bins == kNB && c->geo_feat == kGeo, RF_E_UNSUPPORTED, "%s: n_bins %d / geo_feat_dim %d (16 / 15 are built)", who, c->n_bins, c->geo_feat);
    RF_REQUIRE(c->mlp_precision == 0 || c->mlp_precision == 1, RF_E_UNSUPPORTED, "%s: mlp_precision %d not built (0: fp32 SIMT, 1: tcgen05 bf16x3)", who, c->mlp_precision);
    RF_REQUIRE(c->n_range_d >= 1 && c->n_samples_d >= 0 && c->n_range_d <= kMaxS && c->n_range_d + c->n_samples_d <= kMaxS, RF_E_RANGE,
               "%s: samples per ray %d+%d not in [1,%d]", who, c->n_range_d, c->n_samples_d, kMaxS);
    RF_REQUIRE(n_rays >= 0 && n_rays < (1ll << 40), RF_E_RANGE, "%s: bad ray count", who);
    RF_REQUIRE(c->trunc > 0.f, RF_E_RANGE, "%s: trunc must be positive", who);
    k.n_range_d = c->n_range_d; k.n_samples_d = c->n_samples_d; k.S = c->n_range_d + c->n_samples_d; k.perturb = c->perturb ? 1 : 0;
    k.c_trunc = c->c_trunc; k.trunc = c->trunc; k.clamp_thr = c->clamp_thr; k.clamp_mode = c->clamp_mode ? 1 : 0;
    k.sc_trunc = (float)((double)c->sc_factor * (double)c->trunc);
    k.depth_trunc = c->depth_trunc; k.rgb_all_ones = (c->rgb_missing != 0.f) ? 1 : 0;
    k.n_hash_out = hash->n_levels * 2; k.in1 = k.n_hash_out + kBlob + 1;
    for (int a = 0; a < 3; ++a) { k.b0[a] = c->bbox[2 * a]; k.bl[a] = c->bbox[2 * a + 1] - c->bbox[2 * a]; }
    k.n_rays = n_rays; k.n_total = c->n_rays_total > 0 ? c->n_rays_total : n_rays;
    return 0;
}

template <int HID> static size_t fwd_smem(int in1) { return sizeof(float) * (size_t)(in1 * HID + HID * kOut1 + kIn2 * HID + HID * 4); }
template <int HID> static size_t bwd_smem(int in1) {
    size_t w = in1 * HID + HID * kOut1 + kIn2 * HID + HID * 4;
    size_t dw = HID * in1 + kOut1 * HID + HID * kIn2 + 4 * HID;
    size_t tile = (size_t)kTile * (84 + HID + 20 + HID);
    return sizeof(float) * (w + dw + tile);
}

template <int HID, bool FROM_X>
static int launch_fwd(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o, const float* rays_d,
                      const float* z_vals, const float* xin, long long P, int variant, float* raw, cudaStream_t s) {
    auto fn = sample_fwd_kernel<HID, FROM_X>;
    size_t sm = fwd_smem<HID>(k.in1);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(sample_fwd): %s", cudaGetErrorString(e));
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kTile, sm);
    if (per_sm < 1) per_sm = 1;
    long long tiles = (P + kTile - 1) / kTile;
    int blocks = (int)std::min<long long>(tiles, (long long)num_sms() * per_sm);
    Weights w{p->w_sdf0, p->w_sdf1, p->w_col0, p->w_col1};
    ProfScope ps(RF_PROF_SAMPLE_FWD, s);
    fn<<<blocks, kTile, sm, s>>>(k, hg, gg, p->hash_params, p->gbv_params, w, rays_o, rays_d, z_vals, xin, P, variant, raw);
    RF_CHECK_LAUNCH("sample_fwd_kernel");
    return 0;
}

template <int HID, bool BA>
static int launch_bwd(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o, const float* rays_d,
                      const float* z_vals, long long P, const float* d_raw_tot, const Grads& gr, float* d_pts, cudaStream_t s) {
    auto fn = sample_bwd_kernel<HID, BA>;
    size_t sm = bwd_smem<HID>(k.in1);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(sample_bwd, %zu B): %s", sm, cudaGetErrorString(e));
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kTile, sm);
    if (per_sm < 1) per_sm = 1;
    long long tiles = (P + kTile - 1) / kTile;
    int blocks = (int)std::min<long long>(tiles, (long long)num_sms() * per_sm);
    Weights w{p->w_sdf0, p->w_sdf1, p->w_col0, p->w_col1};
    ProfScope ps(RF_PROF_SAMPLE_BWD, s);
    fn<<<blocks, kTile, sm, s>>>(k, hg, gg, p->hash_params, p->gbv_params, w, rays_o, rays_d, z_vals, P, d_raw_tot, gr, d_pts);
    RF_CHECK_LAUNCH("sample_bwd_kernel");
    return 0;
}

}  // namespace rf

using namespace rf;

extern "C" int rf_ray_sample_z(const rf_ray_cfg* cfg, const float* target_d, const float* u, const float* z_tables,
                               int64_t n_rays, float* z_vals, void* stream) {
    RF_REQUIRE(cfg, RF_E_NULL, "rf_ray_sample_z: NULL cfg");
    RF_REQUIRE(cfg->n_range_d >= 1 && cfg->n_samples_d >= 0 && cfg->n_range_d + cfg->n_samples_d <= kMaxS, RF_E_RANGE, "rf_ray_sample_z: bad sample counts");
    if (n_rays == 0) return 0;
    RF_REQUIRE(target_d && z_tables && z_vals, RF_E_NULL, "rf_ray_sample_z: NULL pointer");
    RF_REQUIRE(!cfg->perturb || u, RF_E_NULL, "rf_ray_sample_z: perturb needs the jitter array u");
    RayK k; memset(&k, 0, sizeof(k));
    k.n_range_d = cfg->n_range_d; k.n_samples_d = cfg->n_samples_d; k.S = k.n_range_d + k.n_samples_d; k.perturb = cfg->perturb ? 1 : 0;
    k.n_rays = n_rays;
    {
        ProfScope ps(RF_PROF_RAY_Z, (cudaStream_t)stream);
        ray_z_kernel<<<(unsigned)((n_rays + 3) / 4), 128, 0, (cudaStream_t)stream>>>(k, target_d, u, z_tables, z_vals);
    }
    RF_CHECK_LAUNCH("ray_z_kernel");
    return 0;
}

extern "C" int rf_ray_composite(const rf_ray_cfg* cfg, const float* raw, const float* z_vals, int64_t n_rays, float* rgb_map,
                                float* depth_map, void* stream) {
    RF_REQUIRE(cfg, RF_E_NULL, "rf_ray_composite: NULL cfg");
    RF_REQUIRE(cfg->n_range_d >= 1 && cfg->n_samples_d >= 0 && cfg->n_range_d + cfg->n_samples_d <= kMaxS, RF_E_RANGE, "rf_ray_composite: bad sample counts");
    RF_REQUIRE(cfg->trunc > 0.f, RF_E_RANGE, "rf_ray_composite: trunc must be positive");
    if (n_rays == 0) return 0;
    RF_REQUIRE(raw && z_vals && rgb_map && depth_map, RF_E_NULL, "rf_ray_composite: NULL pointer");
    RF_REQUIRE(((uintptr_t)raw & 15) == 0, RF_E_ALIGN, "rf_ray_composite: raw must be 16-byte aligned");
    RayK k; memset(&k, 0, sizeof(k));
    k.S = cfg->n_range_d + cfg->n_samples_d; k.trunc = cfg->trunc; k.n_rays = n_rays;
    k.sc_trunc = (float)((double)cfg->sc_factor * (double)cfg->trunc);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(RF_PROF_COMPOSITE_FWD, s);
    launch_composite_fwd(k, raw, z_vals, nullptr, nullptr, rgb_map, depth_map, nullptr, s);
    RF_CHECK_LAUNCH("composite_fwd_kernel");
    return 0;
}

extern "C" int rf_ray_loss_finalize(const double* loss_partials, int64_t n_rays_total, int n_samples, float* losses, void* stream) {
    RF_REQUIRE(loss_partials && losses, RF_E_NULL, "rf_ray_loss_finalize: NULL pointer");
    RF_REQUIRE(n_rays_total > 0 && n_samples > 0, RF_E_RANGE, "rf_ray_loss_finalize: bad counts");
    loss_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(loss_partials, (double)n_rays_total, (double)n_samples, losses);
    RF_CHECK_LAUNCH("loss_finalize_kernel");
    return 0;
}

extern "C" int rf_ray_query_forward(const rf_ray_cfg* cfg, const rf_grid_desc* hash, const rf_grid_desc* gbv, const rf_ray_params* p,
                                    const float* rays_o, const float* rays_d, const float* target_d, const float* target_rgb,
                                    const float* z_vals, int64_t n_rays,
                                    float* raw, float* rgb_map, float* depth_map, double* loss_partials, float* workspace, void* stream) {
    RayK k;
    int rc = make_rayk(k, cfg, hash, gbv, n_rays, "rf_ray_query_forward"); if (rc) return rc;
    if (n_rays == 0) return 0;
    RF_REQUIRE(p && p->hash_params && p->gbv_params && p->w_sdf0 && p->w_sdf1 && p->w_col0 && p->w_col1, RF_E_NULL, "rf_ray_query_forward: NULL parameter pointer");
    RF_REQUIRE(rays_o && rays_d && z_vals && raw && rgb_map && depth_map, RF_E_NULL, "rf_ray_query_forward: NULL pointer");
    RF_REQUIRE(!loss_partials || (target_d && target_rgb), RF_E_NULL, "rf_ray_query_forward: losses need target_d and target_rgb");
    RF_REQUIRE((((uintptr_t)raw | (uintptr_t)p->gbv_params) & 15) == 0 && ((uintptr_t)p->hash_params & 7) == 0, RF_E_ALIGN, "rf_ray_query_forward: raw/gbv need 16-byte, hash 8-byte alignment");
    GridDev hg = to_dev(hash), gg = to_dev(gbv);
    cudaStream_t s = (cudaStream_t)stream;
    long long P = n_rays * k.S;
    if (cfg->mlp_precision == 1 && tc_supported(k, cfg->hidden)) {
        RF_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, RF_E_NULL, "rf_ray_query_forward: mlp_precision 1 needs a 16-byte aligned workspace of rf_ray_workspace_floats()");
        rc = launch_fwd_tc(k, cfg->hidden, hg, gg, p, rays_o, rays_d, z_vals, P, raw, workspace, s);
    } else {
        rc = (cfg->hidden == 64) ? launch_fwd<64, false>(k, hg, gg, p, rays_o, rays_d, z_vals, nullptr, P, 0, raw, s)
                                 : launch_fwd<32, false>(k, hg, gg, p, rays_o, rays_d, z_vals, nullptr, P, 0, raw, s);
    }
    if (rc) return rc;
    {
        ProfScope ps(RF_PROF_COMPOSITE_FWD, s);
        launch_composite_fwd(k, raw, z_vals, target_d, target_rgb, rgb_map, depth_map, loss_partials, s);
    }
    RF_CHECK_LAUNCH("composite_fwd_kernel");
    return 0;
}

extern "C" int rf_ray_query_backward(const rf_ray_cfg* cfg, const rf_grid_desc* hash, const rf_grid_desc* gbv, const rf_ray_params* p,
                                     const float* rays_o, const float* rays_d, const float* target_d, const float* target_rgb, int64_t n_rays,
                                     const float* z_vals, const float* raw, const float* rgb_map, const float* depth_map,
                                     const float* d_rgb_map, const float* d_depth_map, const float* d_raw,
                                     const float* loss_grads, const double* loss_partials,
                                     const rf_ray_grads* g, const float* workspace, float* scratch, void* stream) {
    RayK k;
    int rc = make_rayk(k, cfg, hash, gbv, n_rays, "rf_ray_query_backward"); if (rc) return rc;
    if (n_rays == 0) return 0;
    RF_REQUIRE(p && p->hash_params && p->gbv_params && p->w_sdf0 && p->w_sdf1 && p->w_col0 && p->w_col1, RF_E_NULL, "rf_ray_query_backward: NULL parameter pointer");
    RF_REQUIRE(rays_o && rays_d && z_vals && raw && g && scratch, RF_E_NULL, "rf_ray_query_backward: NULL pointer");
    RF_REQUIRE(!loss_grads || (loss_partials && target_d && target_rgb && rgb_map && depth_map), RF_E_NULL, "rf_ray_query_backward: loss gradients need partials, targets and maps");
    RF_REQUIRE(((uintptr_t)scratch & 15) == 0 && (!g->g_hash || ((uintptr_t)g->g_hash & 7) == 0), RF_E_ALIGN, "rf_ray_query_backward: scratch 16-byte / g_hash 8-byte alignment");
    GridDev hg = to_dev(hash), gg = to_dev(gbv);
    cudaStream_t s = (cudaStream_t)stream;
    long long P = n_rays * k.S;
    const bool tc = cfg->mlp_precision == 1 && tc_supported(k, cfg->hidden);
    float* d_raw_tot = scratch;                 // [P,4]
    int* n_live = tc ? reinterpret_cast<int*>(scratch + 4 * P) : nullptr;   // [N] (tensor-core path), padded to 4
    float* d_pts = scratch + 4 * P;             // [P,3] (BA mode, fp32 SIMT path only)
    {
        ProfScope ps(RF_PROF_COMPOSITE_BWD, s);
        launch_composite_bwd(k, raw, z_vals, rgb_map, depth_map, target_d, target_rgb, d_rgb_map, d_depth_map, d_raw, loss_grads, loss_partials,
                             d_raw_tot, n_live, s);
    }
    RF_CHECK_LAUNCH("composite_bwd_kernel");
    Grads gr{g->g_hash, g->g_w_sdf0, g->g_w_sdf1, g->g_w_col0, g->g_w_col1};
    const bool ba = g->g_rays_o || g->g_rays_d;
    if (tc) {
        RF_REQUIRE(workspace, RF_E_NULL, "rf_ray_query_backward: mlp_precision 1 needs the forward's workspace");
        return launch_bwd_tc(k, cfg->hidden, hg, gg, p, P, workspace, d_raw_tot, n_live, scratch + 4 * P + ((n_rays + 3) & ~3ll), gr,
                             g->g_rays_o, g->g_rays_d, s);
    }
    if (cfg->hidden == 64) rc = ba ? launch_bwd<64, true>(k, hg, gg, p, rays_o, rays_d, z_vals, P, d_raw_tot, gr, d_pts, s)
                                   : launch_bwd<64, false>(k, hg, gg, p, rays_o, rays_d, z_vals, P, d_raw_tot, gr, d_pts, s);
    else rc = ba ? launch_bwd<32, true>(k, hg, gg, p, rays_o, rays_d, z_vals, P, d_raw_tot, gr, d_pts, s)
                 : launch_bwd<32, false>(k, hg, gg, p, rays_o, rays_d, z_vals, P, d_raw_tot, gr, d_pts, s);
    if (rc) return rc;
    if (ba) {
        ray_grad_kernel<<<(unsigned)((n_rays + 3) / 4), 128, 0, s>>>(k, d_pts, z_vals, g->g_rays_o, g->g_rays_d);
        RF_CHECK_LAUNCH("ray_grad_kernel");
    }
    return 0;
}

extern "C" int64_t rf_ray_workspace_floats(const rf_ray_cfg* cfg, const rf_grid_desc* hash, int64_t n_rays) {
    if (!cfg || !hash || cfg->mlp_precision != 1 || n_rays <= 0) return 0;
    return (int64_t)ws_floats((long long)n_rays * (cfg->n_range_d + cfg->n_samples_d), true);
}

extern "C" int64_t rf_point_workspace_floats(const rf_ray_cfg* cfg, const rf_grid_desc* hash, int64_t n) {
    if (!cfg || !hash || cfg->mlp_precision != 1 || n <= 0) return 0;
    return (int64_t)ws_floats((long long)n, false);
}

extern "C" int64_t rf_ray_scratch_floats(const rf_ray_cfg* cfg, const rf_grid_desc* hash, int64_t n_rays, int ray_grads) {
    if (!cfg || !hash || n_rays <= 0) return 0;
    const int64_t P = n_rays * (cfg->n_range_d + cfg->n_samples_d);
    RayK k; k.n_hash_out = hash->n_levels * 2;
    if (cfg->mlp_precision != 1 || !tc_supported(k, cfg->hidden)) return ray_grads ? 7 * P : 4 * P;
    GridDev hg = to_dev(hash);
    return (4 + 2 * hash->n_levels + (ray_grads ? 7 : 0)) * P + (ray_grads ? 4 : 0) + ((n_rays + 3) & ~(int64_t)3) +
           (int64_t)scatter_scratch_floats(hg, (long long)n_rays) + (ws_tiles(P) + 3) / 4 + 4;      // + one byte per tile (liveness flags)
}

extern "C" int rf_point_query_forward(const rf_ray_cfg* cfg, const rf_grid_desc* hash, const rf_grid_desc* gbv, const rf_ray_params* p,
                                      const float* x, int64_t n, int variant, float* raw, float* workspace, void* stream) {
    RayK k;
    int rc = make_rayk(k, cfg, hash, gbv, 0, "rf_point_query_forward"); if (rc) return rc;
    RF_REQUIRE(variant >= 0 && variant <= 2, RF_E_RANGE, "rf_point_query_forward: variant %d", variant);
    RF_REQUIRE(n >= 0, RF_E_RANGE, "rf_point_query_forward: negative n");
    if (n == 0) return 0;
    RF_REQUIRE(p && p->hash_params && p->gbv_params && p->w_sdf0 && p->w_sdf1 && p->w_col0 && p->w_col1 && x && raw, RF_E_NULL, "rf_point_query_forward: NULL pointer");
    RF_REQUIRE(((uintptr_t)raw & 15) == 0, RF_E_ALIGN, "rf_point_query_forward: raw must be 16-byte aligned");
    GridDev hg = to_dev(hash), gg = to_dev(gbv);
    cudaStream_t s = (cudaStream_t)stream;
    k.S = 1;
    if (cfg->mlp_precision == 1 && tc_supported(k, cfg->hidden)) {
        RF_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, RF_E_NULL, "rf_point_query_forward: mlp_precision 1 needs a 16-byte aligned workspace of rf_point_workspace_floats()");
        return launch_points_tc(k, cfg->hidden, hg, gg, p, x, n, variant, raw, workspace, s);
    }
    return (cfg->hidden == 64) ? launch_fwd<64, true>(k, hg, gg, p, nullptr, nullptr, nullptr, x, n, variant, raw, s)
                               : launch_fwd<32, true>(k, hg, gg, p, nullptr, nullptr, nullptr, x, n, variant, raw, s);
}
