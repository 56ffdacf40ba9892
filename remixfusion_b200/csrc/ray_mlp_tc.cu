// ray_mlp_tc.cu — the decoder of the mixed representation on the 5th-generation tensor cores (tcgen05, TMEM).
//
// Replaces ColorSDFNet.forward (model/decoder.py:132-146: SDFNet :59-110, ColorNet :6-53) and its autograd backward,
// including the concatenations of JointEncoding.query_color_sdf (model/scene_rep.py:325-345), for 128-sample tiles:
//
//   reference X1 = [hash32 | oneblob48 | tsdf]           H1 = relu(X1 W0^T)      O = H1 W1^T   (sdf, geo15)
//             X2 = [oneblob48 | geo15 | gbv_rgb3]         H2 = relu(X2 W2^T)      rgb = H2 W3^T
//             raw = (rgb + gbv_rgb, sdf + tsdf)                                   (:344-345)
//   here      the geo features are linear in H1, so they are never materialised: with W21 = W2[:, geo] W1[geo, :] (fp32, per CTA)
//             H2 = relu([oneblob48 | gbv_rgb3] W2'^T + H1 W21^T),   sdf = H1 W1[0]^T
//             — three tensor-core phases per tile forward, five backward (see the kernels), instead of one per layer and
//             direction; the gradients of both factors of W21 come out of M21 = dH2^T H1 when the accumulators are flushed.
//             The weight gradients are accumulated in TMEM across the tiles a CTA processes and handed over with atomics
//             every kFlushTiles tiles.  BA mode additionally produces the gradients w.r.t. the OneBlob inputs and the GBV
//             texel (ray gradients).
//
// Numerics: every GEMM is a bf16x3 product (a_hi b_hi + a_hi b_lo + a_lo b_hi, fp32 accumulation in TMEM), i.e.
// ~2^-16 relative per product — the "rendered colour / depth and gradients within 1e-3" bar with two orders of margin;
// the fp32 SIMT kernels of ray_query.cu (mlp_precision 0) stay as the accuracy anchor.
//
// Structure: a CTA holds the decoder weights once (bf16 hi/lo, chunked no-swizzle layout of umma.cuh) and G
// independent groups of 128 threads (forward) or 256 threads (backward: two threads per tile row, mlp_bwd_tc2_kernel).
// A group owns one tile at a time: thread m stages row m of the operands, one elected thread (elect.sync, operands provably
// warp-uniform: see mlp_bwd_tc2_kernel) issues the MMAs, everybody waits on the group's mbarrier and reads its own TMEM lane.  Groups run out of phase, so one group's tensor-core latency is
// covered by another group's staging.  Inputs are what ray_encode.cu wrote (sample-major: a tile is 128 consecutive rays at
// one sample index): the hash features arrive as READY bf16 hi / lo operand chunks, 16 KB per tile — the forward moves its row
// of them straight into tensor memory, the backward's elected thread fetches the whole block into the X operand with two TMA
// bulk copies (cp.async.bulk -> mbarrier) — plus the GBV features and positions (coalesced, streaming); no random access
// happens here.  Only raw / d_raw ([N][S][4], the reference's layout) are strided.  The backward skips tiles none of whose
// rows has a live upstream gradient (one flag per tile, tile_live_kernel, from the per-ray n_live of composite_bwd_kernel).
#include "ray_common.cuh"
#include "umma.cuh"
#include <stdlib.h>

namespace rf {
namespace {

using namespace umma;

constexpr int kChunkB = 2048;                  // bytes of one 8-column chunk of a 128-row operand
constexpr int kXBlob = 4, kXTail = 10, kXCh = 12;   // X-order chunks: hash 0-3 | oneblob 4-9 | tail 10-11
constexpr int kTailTsdf = 3;                   // tail = [gbv_rgb3 | tsdf | 0 x12]: the geo features never appear as inputs (W21 below)
constexpr int kKX = kXCh * 8;                  // 96
constexpr int kK2 = kKX - 32;                  // colour-net input columns: oneblob 48 | tail 16
constexpr int kFlushTiles = 256;               // weight-gradient accumulators are flushed every this many tiles

template <int HID>
struct WL {                                    // weight shared-memory layout (bytes); B operands, rows = out units
    static constexpr int HC = HID / 8;
    static constexpr int w0 = kXCh * HID * 16; // W0 [HID rows][96]   (X-order columns)
    static constexpr int w1 = HC * 16 * 16;    // W1 [16 rows][HID]
    static constexpr int w2 = 8 * HID * 16;    // W2 [HID rows][64]   (oneblob | tail)
    static constexpr int w3 = HC * 16 * 16;    // W3 [16 rows][HID]   (rows 3..15 zero)
    static constexpr int o_w0h = 0, o_w0l = w0, o_w1h = 2 * w0, o_w1l = o_w1h + w1, o_w2h = o_w1l + w1, o_w2l = o_w2h + w2,
                         o_w3h = o_w2l + w2, o_w3l = o_w3h + w3;
    static constexpr int total = 2 * (w0 + w1 + w2 + w3);
    // backward only: W21 = W2[:, geo] W1[geo, :]  ([HID rows j][HID columns i]), the colour net's view of the sdf net's hidden
    // layer (geo = H1 W1[geo]^T feeds H2 without a nonlinearity in between, model/decoder.py:138-143)
    static constexpr int w21 = HC * HID * 16;
    static constexpr int o_w21h = total, o_w21l = total + w21;
    static constexpr int total_bwd = total + 2 * w21;       // (the forward carries W21 too)
};

__device__ __forceinline__ void store_split(unsigned char* hi, unsigned char* lo, int rows, int r, int kx, float v) {
    __nv_bfloat16 h, l; split_bf16(v, h, l);
    uint32_t off = chunk_off(rows, r, kx >> 3) + (uint32_t)(kx & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo + off) = l;
}

// nn.Linear weights ([out][in], model/decoder.py:31-47,87-102) -> chunked bf16 hi/lo B operands
template <int HID>
__device__ void load_weights(unsigned char* w, const Weights& wt, int nthreads) {
    using L = WL<HID>;
    constexpr int in1 = 81;
    for (int i = threadIdx.x; i < HID * kKX; i += nthreads) {
        int j = i / kKX, kx = i - j * kKX;
        float v = 0.f;
        if (kx < 32) v = wt.w_sdf0[j * in1 + hash_col_to_feature(kx)];      // X-order of the hash columns (ray_common.cuh)
        else if (kx < 80) v = wt.w_sdf0[j * in1 + kx];
        else if (kx == 80 + kTailTsdf) v = wt.w_sdf0[j * in1 + 80];
        store_split(w + L::o_w0h, w + L::o_w0l, HID, j, kx, v);
    }
    for (int i = threadIdx.x; i < 16 * HID; i += nthreads) {
        int r = i / HID, j = i - r * HID;
        store_split(w + L::o_w1h, w + L::o_w1l, 16, r, j, wt.w_sdf1[r * HID + j]);
        store_split(w + L::o_w3h, w + L::o_w3l, 16, r, j, (r < 3) ? wt.w_col1[r * HID + j] : 0.f);
    }
    for (int i = threadIdx.x; i < HID * kK2; i += nthreads) {                  // colour net: blob | gbv rgb (its geo columns live in W21)
        int j = i / kK2, kx = i - j * kK2;
        float v = 0.f;
        if (kx < kBlob) v = wt.w_col0[j * kIn2 + kx];
        else if (kx < kBlob + 3) v = wt.w_col0[j * kIn2 + kGeo + kx];
        store_split(w + L::o_w2h, w + L::o_w2l, HID, j, kx, v);
    }
}

// W21[j][i] = sum_g W2[j][48 + g] W1[1 + g][i] (fp32), stored like the other weights; w1s = W1[0][:] (the sdf row) as plain fp32
template <int HID>
__device__ void load_w21(unsigned char* w, float* w1s, const Weights& wt, int nthreads) {
    using L = WL<HID>;
    for (int i = threadIdx.x; i < HID * HID; i += nthreads) {
        const int j = i / HID, c = i - j * HID;
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < kGeo; ++g) v = fmaf(wt.w_col0[j * kIn2 + kBlob + g], wt.w_sdf1[(1 + g) * HID + c], v);
        store_split(w + L::o_w21h, w + L::o_w21l, HID, j, c, v);
    }
    for (int i = threadIdx.x; i < HID; i += nthreads) w1s[i] = wt.w_sdf1[i];
}

__device__ __forceinline__ void grp_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ void grp_wait(uint64_t* bar, uint32_t& phase) {
    if (!mbar_wait(bar, phase)) __trap();          // a tensor-core pipeline that never completes must fail loudly
    phase ^= 1u;
    fence_after_sync();
}

// row m, chunk c of a 128-row operand <- 8 floats (hi and lo parts)
__device__ __forceinline__ void stage8(unsigned char* hi, unsigned char* lo, int m, int c, const float* v) {
    uint4 h, l; split8(v, h, l);
    uint32_t off = chunk_off(128, m, c);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
}
__device__ __forceinline__ void stage_zero(unsigned char* hi, unsigned char* lo, int m, int c) {
    uint32_t off = chunk_off(128, m, c);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(0, 0, 0, 0);
}

__device__ __forceinline__ float cdf_u(float u) {
    float u2 = u * u, u4 = u2 * u2;
    return fminf(fmaxf((15.0f / 16.0f) * u * (1.0f - (2.0f / 3.0f) * u2 + (1.0f / 5.0f) * u4) + 0.5f, 0.0f), 1.0f);
}
// OneBlob of one coordinate (model/encodings.py:65-76; Appendix B7) into chunks c, c+1 of row m.  The quartic kernel
// has support +-1/16, so only the bin holding x and its two neighbours are non-zero: inside [-0.9, 1.9] (where the
// three periodic copies of Appendix B7 cover the support) the 16 bins are three values; elsewhere the general form.
__device__ __forceinline__ void stage_oneblob(float x, unsigned char* hi, unsigned char* lo, int m, int c) {
    if (x > -0.9f && x < 1.9f) {
        stage_zero(hi, lo, m, c); stage_zero(hi, lo, m, c + 1);
        float xw = x - floorf(x);
        if (xw >= 1.0f) xw = 0.0f;
        float t = xw * (float)kNB;
        int b = (int)t;
        float ub = (float)b - t;
        float cb = cdf_u(ub), cb1 = cdf_u(ub + 1.0f);
        const int km = (b - 1) & (kNB - 1), kp = (b + 1) & (kNB - 1);
        __nv_bfloat16 h, l;
        split_bf16(cb, h, l);
        uint32_t o = chunk_off(128, m, c + (km >> 3)) + (uint32_t)(km & 7) * 2u;
        *reinterpret_cast<__nv_bfloat16*>(hi + o) = h; *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
        split_bf16(cb1 - cb, h, l);
        o = chunk_off(128, m, c + (b >> 3)) + (uint32_t)(b & 7) * 2u;
        *reinterpret_cast<__nv_bfloat16*>(hi + o) = h; *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
        split_bf16(1.0f - cb1, h, l);
        o = chunk_off(128, m, c + (kp >> 3)) + (uint32_t)(kp & 7) * 2u;
        *reinterpret_cast<__nv_bfloat16*>(hi + o) = h; *reinterpret_cast<__nv_bfloat16*>(lo + o) = l;
    } else {
        float ob[kNB];
        oneblob_coord<kNB>(x, ob);
        stage8(hi, lo, m, c, ob); stage8(hi, lo, m, c + 1, ob + 8);
    }
}

// ---- MMA issue helpers (one thread).  All operands are bf16 hi/lo pairs; each k-step issues the bf16x3 triple. ----
// The descriptors of consecutive k-steps differ only in the start-address field (bits 0..13, 16-byte units), so they are
// built once and advanced with one 64-bit add per operand (shared memory is < 256 KB: no carry out of the field).
template <int NKS>
__device__ __forceinline__ void mma_steps(uint32_t d, uint64_t dah, uint64_t dal, uint64_t dbh, uint64_t dbl, uint32_t a_step, uint32_t b_step,
                                          uint32_t idesc, uint32_t& acc) {
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
        mma_bf16(d, dah, dbh, idesc, acc); acc = 1;
        mma_bf16(d, dah, dbl, idesc, 1);
        mma_bf16(d, dal, dbh, idesc, 1);
        dah = desc_advance(dah, a_step); dal = desc_advance(dal, a_step); dbh = desc_advance(dbh, b_step); dbl = desc_advance(dbl, b_step);
    }
}
// A [128 x K] K-major at consecutive chunks; B = weights [brows = N rows] K-major, chunks from the given address
template <int NKS>
__device__ __forceinline__ void mma_kk(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, int brows, uint32_t idesc, uint32_t& acc) {
    mma_steps<NKS>(d, smem_desc(ah, kChunkB, 128), smem_desc(al, kChunkB, 128), smem_desc(bh, brows * 16, 128), smem_desc(bl, brows * 16, 128),
                   2 * kChunkB, 2 * brows * 16, idesc, acc);
}
// A [128 x K] K-major; B = weights stored [brows = K rows][chunks over N], used MN-major from chunk address bh/bl
template <int NKS>
__device__ __forceinline__ void mma_km(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, int brows, uint32_t idesc, uint32_t& acc) {
    mma_steps<NKS>(d, smem_desc(ah, kChunkB, 128), smem_desc(al, kChunkB, 128), smem_desc(bh, 128, brows * 16), smem_desc(bl, 128, brows * 16),
                   2 * kChunkB, 2 * 128, idesc, acc);
}
// Weight gradients D[f][j] (+)= sum over the 128 samples of A[m][f] B[m][j]: both operands [128 rows] used MN-major.
// ONE MMA per k-step: `a` and `b` each address a hi block immediately
// followed by its lo block.  M = 128 then spans [A_hi features | A_lo features | ...] and N spans [B_hi | B_lo], so
// the accumulator holds hi*hi (rows < F, cols < n), hi*lo (rows < F, cols >= n) and lo*hi (rows >= F, cols < n) at once.
__device__ __forceinline__ void mma_mm1(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t& acc) {
    uint64_t da = smem_desc(a, 128, kChunkB), db = smem_desc(b, 128, kChunkB);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        mma_bf16(d, da, db, idesc, acc); acc = 1;
        da = desc_advance(da, 2 * 128); db = desc_advance(db, 2 * 128);
    }
}
struct TileIn { uint4 hh[4], hl[4]; float4 g; float x[3]; };   // row m: 4 hash operand chunks (hi, lo), GBV texel, position

// Plane index q = s * N + r of the first row of a group's current tile, kept as (s, r) and advanced by a fixed number of
// rows per iteration: no 64-bit division per thread per tile.
struct TilePos {
    long long N; int S;
    long long s, r;            // of row 0 of the tile
    long long ds, dr;          // advance per iteration
    __device__ void init(long long n_rays, int S_, long long q0, long long step) {
        N = n_rays; S = S_; s = q0 / N; r = q0 - s * N; ds = step / N; dr = step - ds * N;
    }
    __device__ void next() { s += ds; r += dr; if (r >= N) { r -= N; ++s; } }
    // (sample, ray) of row m (m < 128; N may be smaller than 128)
    __device__ void row(int m, long long& ss, long long& rr) const {
        rr = r + m; ss = s;
        while (rr >= N) { rr -= N; ++ss; }
    }
    // raw index r * S + s of row m
    __device__ long long raw_index(int m) const {
        long long rr, ss; row(m, ss, rr);
        return rr * S + ss;
    }
};

// row m of tile `tile` (plane index q = tile * 128 + m).  The operand block of the last tile is zero-filled beyond P.
__device__ __forceinline__ void load_tile(TileIn& t, const float* __restrict__ feat, long long P, long long tile, int m) {
    const long long q = tile * kTile + m;
    if (tile * kTile < P) {
        const uint4* hop = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(feat) + tile * kHopTileBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) { t.hh[c] = __ldg(hop + c * 128 + m); t.hl[c] = __ldg(hop + 512 + c * 128 + m); }
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) { t.hh[c] = make_uint4(0, 0, 0, 0); t.hl[c] = make_uint4(0, 0, 0, 0); }
    }
    if (q < P) {
        t.g = __ldg(reinterpret_cast<const float4*>(feat + ws_off_gbv(P)) + q);
        const float* xn = feat + ws_off_xn(P);
        t.x[0] = __ldg(xn + q); t.x[1] = __ldg(xn + P + q); t.x[2] = __ldg(xn + 2 * P + q);
    } else {
        t.g = make_float4(0.f, 0.f, 0.f, 0.f);
        t.x[0] = t.x[1] = t.x[2] = 0.5f;
    }
}

// Pull a later tile of this group towards L2 while the current one is processed: the operand block is 128 lines of 128 B
// (one per thread), GBV 16 lines, xn 3 x 4 lines.
__device__ __forceinline__ void prefetch_tile(const float* __restrict__ feat, long long P, long long tile, int m, bool with_hash) {
    const long long q0 = tile * kTile;
    if (q0 >= P) return;
    const long long last = P - 1;
    if (with_hash) prefetch_l2(reinterpret_cast<const unsigned char*>(feat) + tile * kHopTileBytes + m * 128);
    if (m < 16) prefetch_l2(feat + ws_off_gbv(P) + 4 * min(q0 + 8 * m, last));
    else if (m < 28) prefetch_l2(feat + ws_off_xn(P) + (long long)((m - 16) >> 2) * P + min(q0 + 32 * ((m - 16) & 3), last));
}

// ------------------------------------------------------------------------------------------------------------
// Forward.  The decoder kernels are bound by shared-memory bandwidth (tensor-core operand reads + staging stores) and by the
// serial staging -> MMA -> read-back chain of a tile, so the forward keeps every A operand it can in TENSOR MEMORY (a thread
// writes its row of the hash features, of the tail and of the hidden activations as bf16 pairs with tcgen05.st and the MMAs
// read them with the TS form; only the OneBlob block stays in shared memory: three scalars per coordinate at data-dependent
// columns) and runs THREE phases per tile instead of one per layer: the geo features are linear in H1 (model/decoder.py:138-143),
// so the colour net's hidden layer is issued together with the sdf output, from H1 and W21 = W2[:, geo] W1[geo, :]:
//   1  H1 = relu(X1 W0^T)      2  sdf = H1 W1[0]^T,  H2 = relu(X2' W2^T + H1 W21^T)      3  rgb = H2 W3^T
// TMEM columns of a group: accumulator [0, HID) | hash hi/lo 16+16, later the hidden activations hi/lo HID/2 + HID/2 over the same
// columns (the hash operand is dead once phase 1 has completed, H1 once phase 2 has) | tail hi/lo 8+8 | sdf result 16:
// 96 columns at hidden 32 (four groups per SM; five fit but are slower, see launch_fwd_tc), 160 at hidden 64 (three).
// ------------------------------------------------------------------------------------------------------------
template <int HID>
struct FwdL {
    static constexpr int c_blob_hi = 0, c_blob_lo = 6, chunks = 12;     // shared memory per group: OneBlob hi / lo
    static constexpr int bytes = chunks * kChunkB;
    static constexpr int OPW = HID > 32 ? HID : 32;                     // operand block: hash (32 columns) or hidden (HID)
    static constexpr int t_acc = 0, t_hash_hi = HID, t_hash_lo = HID + 16, t_h_hi = HID, t_h_lo = HID + HID / 2,
                         t_tail_hi = HID + OPW, t_tail_lo = t_tail_hi + 8, t_o = t_tail_hi + 16, tcols = t_o + 16;
};

// A from TMEM (hi at column ah, lo at column al, 8 columns per k-step); B = weights K-major
template <int NKS>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, int brows, uint32_t idesc, uint32_t& acc) {
    uint64_t dbh = smem_desc(bh, brows * 16, 128), dbl = smem_desc(bl, brows * 16, 128);
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
        mma_bf16_ts(d, ah + 8 * s, dbh, idesc, acc); acc = 1;
        mma_bf16_ts(d, ah + 8 * s, dbl, idesc, 1);
        mma_bf16_ts(d, al + 8 * s, dbh, idesc, 1);
        dbh = desc_advance(dbh, 2 * brows * 16); dbl = desc_advance(dbl, 2 * brows * 16);
    }
}
// 8 floats -> bf16 hi / lo pairs -> 4 TMEM columns each (chunk c of an operand whose hi / lo blocks start at th / tl)
__device__ __forceinline__ void tstage8(uint32_t th, uint32_t tl, int c, const float* v) {
    uint4 h, l; split8(v, h, l);
    tmem_st4(th + 4 * c, h);
    tmem_st4(tl + 4 * c, l);
}
__device__ __forceinline__ void tstage_zero(uint32_t th, uint32_t tl, int c) {
    tmem_st4(th + 4 * c, make_uint4(0, 0, 0, 0));
    tmem_st4(tl + 4 * c, make_uint4(0, 0, 0, 0));
}
template <int HID>
__device__ __forceinline__ void relu_to_tmem(uint32_t tacc, uint32_t th, uint32_t tl) {
#pragma unroll
    for (int q = 0; q < HID / 16; ++q) {                      // 16 columns at a time (register pressure)
        float v[16];
        tmem_ld16(tacc + 16 * q, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        tstage8(th, tl, 2 * q, v); tstage8(th, tl, 2 * q + 1, v + 8);
    }
}

template <int HID, int G>
__global__ void __launch_bounds__(G * 128, 1) mlp_fwd_tc_kernel(RayK k, Weights wts, const float* __restrict__ feat, long long P,
                                                                int variant, float* __restrict__ raw) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[G];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float w1s[HID];
    using W = WL<HID>; using A = FwdL<HID>;
    constexpr int HC = HID / 8;
    constexpr uint32_t TCOLS = (G * A::tcols <= 128) ? 128 : (G * A::tcols <= 256) ? 256 : 512;
    static_assert(G * A::tcols <= 512, "TMEM columns");
    const int tid = threadIdx.x, m = tid & 127, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);                  // warp-uniform for the compiler: MMA operands in uniform registers
    const int g = warp_u >> 2;
    const bool iwarp = (warp_u & 3) == 0;                                  // the group's MMA-issuing warp (uniform)
    unsigned char* wsm = smem + G * A::bytes;
    unsigned char* act = smem + g * A::bytes;
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 0) { for (int i = 0; i < G; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    load_weights<HID>(wsm, wts, G * 128);
    load_w21<HID>(wsm, w1s, wts, G * 128);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0) + (uint32_t)g * A::tcols;   // group's columns (lane 0: MMA operand addresses)
    const uint32_t tlane = tb + ((uint32_t)((warp & 3) * 32) << 16);       // this thread's lane quadrant
    uint64_t* bar = &bars[g];
    uint32_t phase = 0;
    unsigned char *blob_hi = act + A::c_blob_hi * kChunkB, *blob_lo = act + A::c_blob_lo * kChunkB;
    const uint32_t w0h = smem_u32(wsm + W::o_w0h), w0l = smem_u32(wsm + W::o_w0l), w1h = smem_u32(wsm + W::o_w1h), w1l = smem_u32(wsm + W::o_w1l);
    const uint32_t w2h = smem_u32(wsm + W::o_w2h), w2l = smem_u32(wsm + W::o_w2l), w3h = smem_u32(wsm + W::o_w3h), w3l = smem_u32(wsm + W::o_w3l);
    const uint32_t w21h = smem_u32(wsm + W::o_w21h), w21l = smem_u32(wsm + W::o_w21l);
    constexpr uint32_t idH = idesc_bf16(HID, false, false), id16 = idesc_bf16(16, false, false);

    // Inputs are software-pipelined through their own registers: a tile's features and positions are dead once its X row
    // is staged, so the same registers are reloaded with the NEXT tile's inputs right away and those loads are in flight
    // during the MMA phases (the first use of freshly loaded inputs was ~20 % of all stall samples).
    const long long tstep = (long long)gridDim.x * G;
    TilePos tp; tp.init(k.n_rays, k.S, ((long long)blockIdx.x * G + g) * kTile, tstep * kTile);
    TileIn t;
    load_tile(t, feat, P, (long long)blockIdx.x * G + g, m);
    for (long long tile = (long long)blockIdx.x * G + g; tile * kTile < P; tile += tstep, tp.next()) {
        const long long q = tile * kTile + m;                                                 // plane index s * N + r
        const bool live = q < P;
        const long long p = live ? tp.raw_index(m) : 0;                                       // raw index r * S + s
        prefetch_tile(feat, P, tile + 2 * tstep, m, true);
        const float4 gb = t.g;                                                                // GBV features of this tile
        float t_add, cin, d0, d1;
        tsdf_terms(k, variant, gb.x, t_add, cin, d0, d1);                                     // scene_rep.py:330-337 (:230-233, :292-294)
        // X row: hash (ready bf16 hi / lo chunks) -> TMEM, OneBlob -> shared memory, tail = [gbv rgb | decoder tsdf input | 0 x12] -> TMEM
#pragma unroll
        for (int c = 0; c < 4; ++c) { tmem_st4(tlane + A::t_hash_hi + 4 * c, t.hh[c]); tmem_st4(tlane + A::t_hash_lo + 4 * c, t.hl[c]); }
        if (live) {
#pragma unroll
            for (int a = 0; a < 3; ++a) stage_oneblob(t.x[a], blob_hi, blob_lo, m, 2 * a);
        } else {
            for (int c = 0; c < 6; ++c) stage_zero(blob_hi, blob_lo, m, c);
        }
        {
            float v1[8] = {gb.y, gb.z, gb.w, cin, 0.f, 0.f, 0.f, 0.f};
            tstage8(tlane + A::t_tail_hi, tlane + A::t_tail_lo, 0, v1);
            tstage_zero(tlane + A::t_tail_hi, tlane + A::t_tail_lo, 1);
        }
        load_tile(t, feat, P, tile + tstep, m);                                               // next tile's inputs (see above)
        tmem_st_wait();
        fence_async_smem(); fence_before_sync(); grp_sync(g);
        if (iwarp && elect_one()) {                                                                         // 1: H1 = X1 W0^T
            fence_after_sync();
            uint32_t acc = 0;
            mma_ts<2>(tb + A::t_acc, tb + A::t_hash_hi, tb + A::t_hash_lo, w0h, w0l, HID, idH, acc);
            mma_kk<3>(tb + A::t_acc, smem_u32(blob_hi), smem_u32(blob_lo), w0h + kXBlob * HID * 16, w0l + kXBlob * HID * 16, HID, idH, acc);
            mma_ts<1>(tb + A::t_acc, tb + A::t_tail_hi, tb + A::t_tail_lo, w0h + kXTail * HID * 16, w0l + kXTail * HID * 16, HID, idH, acc);
            commit(bar);
        }
        grp_wait(bar, phase);
        relu_to_tmem<HID>(tlane + A::t_acc, tlane + A::t_h_hi, tlane + A::t_h_lo);            // decoder.py:105-107
        tmem_st_wait();
        fence_before_sync(); grp_sync(g);
        if (iwarp && elect_one()) {                                                                         // 2: O = H1 W1^T, H2 = X2' W2^T + H1 W21^T
            fence_after_sync();
            uint32_t acc = 0;
            mma_ts<HC / 2>(tb + A::t_o, tb + A::t_h_hi, tb + A::t_h_lo, w1h, w1l, 16, id16, acc);
            acc = 0;
            mma_kk<3>(tb + A::t_acc, smem_u32(blob_hi), smem_u32(blob_lo), w2h, w2l, HID, idH, acc);
            mma_ts<1>(tb + A::t_acc, tb + A::t_tail_hi, tb + A::t_tail_lo, w2h + 6 * HID * 16, w2l + 6 * HID * 16, HID, idH, acc);
            mma_ts<HC / 2>(tb + A::t_acc, tb + A::t_h_hi, tb + A::t_h_lo, w21h, w21l, HID, idH, acc);
            commit(bar);
        }
        grp_wait(bar, phase);
        float sdf;
        {
            float o16[16];
            tmem_ld16(tlane + A::t_o, o16);
            sdf = o16[0] + t_add;                                                             // scene_rep.py:345
        }
        relu_to_tmem<HID>(tlane + A::t_acc, tlane + A::t_h_hi, tlane + A::t_h_lo);            // decoder.py:49-51
        tmem_st_wait();
        fence_before_sync(); grp_sync(g);
        if (iwarp && elect_one()) {                                                                         // 3: rgb = H2 W3^T
            fence_after_sync();
            uint32_t acc = 0;
            mma_ts<HC / 2>(tb + A::t_acc, tb + A::t_h_hi, tb + A::t_h_lo, w3h, w3l, 16, id16, acc);
            commit(bar);
        }
        grp_wait(bar, phase);
        {
            float o16[16];
            tmem_ld16(tlane + A::t_acc, o16);
            if (live) reinterpret_cast<float4*>(raw)[p] = make_float4(o16[0] + gb.y, o16[1] + gb.z, o16[2] + gb.w, sdf);   // :344-345
        }
        fence_before_sync();
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, TCOLS);
}

// ------------------------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------------------------
template <int HID>
struct BwdL {
    static constexpr int HC = HID / 8;
    // group region (chunks): X hi [hash4|blob6|tail2] | X lo 12 | H1 hi,lo | H2 hi,lo | D hi 1 | D lo 1
    static constexpr int c_x_hi = 0, c_x_lo = kXCh, c_h1_hi = 2 * kXCh, c_h1_lo = c_h1_hi + HC, c_h2_hi = c_h1_lo + HC, c_h2_lo = c_h2_hi + HC,
                         c_d_hi = c_h2_lo + HC, c_d_lo = c_d_hi + 1;
    static constexpr int chunks = c_d_lo + 1;
    static constexpr int bytes = chunks * kChunkB;
    // TMEM columns of a group.  Weight-gradient accumulators are transposed (lane = a row of the MN-major A window, column = a
    // column of the MN-major B window):
    //   t_w0 (96)   lanes [dH1 hi | dH1 lo | dH2 hi | dH2 lo] x 32 units (hidden 32; hidden 64: [dH1 hi | dH1 lo]), columns = X-order features
    //   t_w2 (64)   hidden 64 only: lanes [dH2 hi | dH2 lo], columns = blob | tail
    //   t_w3 (16)   lanes [H2 hi | H2 lo | ...], columns [dRGB hi 8 | dRGB lo 8]
    //   t_m (2 HID) lanes [dH2 hi | dH2 lo | D hi 8 | D lo 8 | ...], columns [H1 hi | H1 lo]:  M21 = dH2^T H1 and (hidden 32) dsdf^T H1
    //   t_w1 (16)   hidden 64 only: lanes [H1 hi | H1 lo], columns [dsdf hi 8 | dsdf lo 8]
    //   t_a         the tile's working accumulator (HID columns; 64 in BA mode, over the slot), t_s the K-major operand slot [hi HID/2 | lo HID/2]
    static constexpr int t_w0 = 0, t_w3 = 96, t_w2 = 128, t_w1 = 208,
                         t_m = (HID == 32) ? 128 : 256, t_a = (HID == 32) ? 192 : 384, t_s = t_a + ((HID == 32) ? 32 : 64);
    static constexpr int tcols = (HID == 32) ? 256 : 512;
};

// ------------------------------------------------------------------------------------------------------------
// Backward, two threads per tile row: a group is 256 threads and thread (m, h) handles half h of everything row m stages or
// reads back (hidden units [h HID/2, (h+1) HID/2), ...), so the serial SIMT stretch between two tensor-core phases is half as
// long and a CTA runs 16 warps (hidden 32).
//
// FIVE tensor-core phases per tile.  The geo features are linear in H1 (O = H1 W1^T, no activation, model/decoder.py:138-143), so
// the colour net's hidden layer is computed from H1 directly with the combined matrix W21 = W2[:, geo] W1[geo, :], and its
// gradient flows back the same way — neither O nor d geo is ever materialised:
//   1  H1 = relu(X1 W0^T)
//   2  H2 = relu(X2' W2^T + H1 W21^T)                       X2' = [blob | 0 x15 | gbv rgb]
//   3  dH2 = (dRGB W3) . relu'                               dW3^T += H2^T dRGB
//   4  dH1 = (dH2 W21 + dsdf W1[0]) . relu'                  M21 += dH2^T H1,  dW1[0] += dsdf^T H1   (hidden 64: dW2 += dH2^T X2')
//   5  d hash = dH1 W0[:, hash]                              dW0 += dH1^T X1  (hidden 32: and dW2 += dH2^T X2' in the same MMAs)
// When the accumulators are flushed, the gradients of the two factors of W21 come out of M21 (HID x HID, fp32):
//   dW2[:, geo] = M21 W1[geo]^T,   dW1[geo] = W2[:, geo]^T M21.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grp_sync2(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

// row m, chunk c of a narrow operand that is read both ways: MN-major from shared memory (weight gradients) and as the
// K-major A operand of the next GEMM from tensor memory (hi block at th, lo block at tl, 4 columns per chunk)
__device__ __forceinline__ void stage8_both(unsigned char* hi, unsigned char* lo, int m, int c, const float* v, uint32_t th, uint32_t tl) {
    uint4 h, l; split8(v, h, l);
    uint32_t off = chunk_off(128, m, c);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
    tmem_st4(th + 4 * c, h);
    tmem_st4(tl + 4 * c, l);
}
// A from TMEM; B = weights stored [brows = K rows][chunks over N], used MN-major (cf. mma_km)
template <int NKS>
__device__ __forceinline__ void mma_ts_m(uint32_t d, uint32_t ah, uint32_t al, uint32_t bh, uint32_t bl, int brows, uint32_t idesc, uint32_t& acc) {
    uint64_t dbh = smem_desc(bh, 128, brows * 16), dbl = smem_desc(bl, 128, brows * 16);
#pragma unroll
    for (int s = 0; s < NKS; ++s) {
        mma_bf16_ts(d, ah + 8 * s, dbh, idesc, acc); acc = 1;
        mma_bf16_ts(d, ah + 8 * s, dbl, idesc, 1);
        mma_bf16_ts(d, al + 8 * s, dbh, idesc, 1);
        dbh = desc_advance(dbh, 2 * 128); dbl = desc_advance(dbl, 2 * 128);
    }
}

template <int NC>      // NC (16 | 32) accumulator columns of this lane -> relu -> chunks c0.. ; returns the relu mask
__device__ __forceinline__ uint32_t relu_half(uint32_t taddr, unsigned char* hi, unsigned char* lo, int m, int c0, uint32_t th, uint32_t tl) {
    uint32_t mk = 0;
#pragma unroll
    for (int q = 0; q < NC / 16; ++q) {
        float v[16];
        tmem_ld16(taddr + 16 * q, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) { if (v[i] > 0.f) mk |= (1u << (16 * q + i)); v[i] = fmaxf(v[i], 0.f); }
        stage8_both(hi, lo, m, c0 + 2 * q, v, th, tl); stage8_both(hi, lo, m, c0 + 2 * q + 1, v + 8, th, tl);
    }
    return mk;
}
// (accumulator + a * w[.]) . mask -> chunks c0.. (shared memory and operand slot); w: NC fp32 values in shared memory (broadcast reads)
template <int NC>
__device__ __forceinline__ void masked_half(uint32_t taddr, unsigned char* hi, unsigned char* lo, int m, int c0, uint32_t mask, uint32_t th, uint32_t tl,
                                            float a = 0.f, const float* w = nullptr) {
#pragma unroll
    for (int q = 0; q < NC / 16; ++q) {
        float v[16];
        tmem_ld16(taddr + 16 * q, v);
        if (w) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 ww = *reinterpret_cast<const float4*>(w + 16 * q + i);
                v[i] = fmaf(a, ww.x, v[i]); v[i + 1] = fmaf(a, ww.y, v[i + 1]); v[i + 2] = fmaf(a, ww.z, v[i + 2]); v[i + 3] = fmaf(a, ww.w, v[i + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ((mask >> (16 * q + i)) & 1u) ? v[i] : 0.f;
        stage8_both(hi, lo, m, c0 + 2 * q, v, th, tl); stage8_both(hi, lo, m, c0 + 2 * q + 1, v + 8, th, tl);
    }
}

struct TileHalf { float4 g; float x0, x1, xb; uint4 hh[2], hl[2]; };   // h = 0: x, y; h = 1: z, GBV texel; hash chunks 2h, 2h + 1 (register path)

// xb (BA mode): the second coordinate whose OneBlob gradient this thread reduces: h = 0 handles x then z, h = 1 handles
// y then the GBV texel gradient.  TMAH: the hash chunks arrive by TMA bulk copy instead of through this thread's registers.
template <bool BA, bool TMAH>
__device__ __forceinline__ void load_half(TileHalf& t, const float* __restrict__ feat, long long P, long long tile, int m, int h) {
    const long long q = tile * kTile + m;
    t.g = make_float4(0.f, 0.f, 0.f, 0.f); t.x0 = t.x1 = t.xb = 0.5f;
    if (!TMAH) {                                              // the operand block of the last tile is zero-filled beyond P
        const uint4* hop = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(feat) + tile * kHopTileBytes) + m;
#pragma unroll
        for (int c = 0; c < 2; ++c) { t.hh[c] = __ldg(hop + (2 * h + c) * 128); t.hl[c] = __ldg(hop + 512 + (2 * h + c) * 128); }
    }
    if (q < P) {
        const float* xn = feat + ws_off_xn(P);
        if (h == 0) { t.x0 = __ldg(xn + q); t.x1 = __ldg(xn + P + q); if (BA) t.xb = __ldg(xn + 2 * P + q); }
        else { t.x0 = __ldg(xn + 2 * P + q); t.g = __ldg(reinterpret_cast<const float4*>(feat + ws_off_gbv(P)) + q); if (BA) t.xb = __ldg(xn + P + q); }
    }
}

// Input-layer weight gradients with the hidden-gradient block as the A operand (MN-major, M = 128 rows = 16 consecutive
// chunks: [dH hi | dH lo] of one net at hidden 64, of both nets at hidden 32) and X as the B operand (MN-major, N = 8 per
// chunk), first its hi part then its lo part into the SAME columns: D[row][f] += sum_m dH_part[m][row] (X_hi + X_lo)[m][f].
// The hi and lo rows of a unit are added when the accumulators are flushed.  Compared with X as the A operand this reads
// 15 KB instead of 22 KB of operands per k-step for both nets and needs half the MMAs.
__device__ __forceinline__ void mma_wx(uint32_t d, uint32_t a, uint32_t bh, uint32_t bl, uint32_t idesc, uint32_t& acc) {
    uint64_t da = smem_desc(a, 128, kChunkB), dbh = smem_desc(bh, 128, kChunkB), dbl = smem_desc(bl, 128, kChunkB);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        mma_bf16(d, da, dbh, idesc, acc); acc = 1;
        mma_bf16(d, da, dbl, idesc, 1);
        da = desc_advance(da, 2 * 128); dbh = desc_advance(dbh, 2 * 128); dbl = desc_advance(dbl, 2 * 128);
    }
}

// Flush of the weight-gradient accumulators (layouts: BwdL).  `scratch`: HID * HID floats of the group's shared memory (the
// H2 blocks, idle between tiles) for M21; every thread of the group calls this.
template <int HID>
__device__ __forceinline__ void flush_wgrads3(uint32_t tlane, const Grads& gr, const Weights& wt, float* scratch, int g, int m, int h) {
    using A = BwdL<HID>;
    fence_after_sync();
    {   // input layers: dW0 (sdf net) and dW2 (colour net) from the X products
        const int j = m & (HID - 1);
        const bool colour = (HID == 32) ? (m >= 64) : (h == 1);
        float* gx = colour ? gr.g_w_col0 : gr.g_w_sdf0;
        const int ld = colour ? kIn2 : 81;
        const uint32_t tx = tlane + ((HID == 64 && h == 1) ? A::t_w2 : A::t_w0);
        const int q0 = (HID == 32) ? (h ? 3 : 0) : 0;
        const int q1 = (HID == 32) ? (h ? 6 : 3) : (h ? 4 : 6);
        const int fshift = (HID == 32 && colour) ? -32 : 0;       // X-order column -> colour-net input (blob | tail)
#pragma unroll 1
        for (int q = q0; q < q1; ++q) {
            float v[16];
            tmem_ld16(tx + 16 * q, v);
            if (!gx) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int f = 16 * q + i + fshift;
                int c = -1;
                if (colour) { if (f >= 0 && f < kBlob) c = f; else if (f >= kBlob && f < kBlob + 3) c = f + kGeo; }     // (the geo columns come from M21 below)
                else { if (f < 32) c = hash_col_to_feature(f); else if (f < 80) c = f; else if (f == 80 + kTailTsdf) c = 80; }
                if (c >= 0) atomicAdd(gx + j * ld + c, v[i]);
            }
        }
    }
    // the products below are [hi rows | lo rows] x [hi columns | lo columns]: value = hi.hi + hi.lo + lo.hi
    const bool hi_lane = m < HID, lo_lane = m >= HID && m < 2 * HID;
    const int jj = m & (HID - 1);
    if (h == 1) {
        float v[16];                                              // dW3[c][j], c < 3 (model/decoder.py:44-47)
        tmem_ld16(tlane + A::t_w3, v);
        if (gr.g_w_col1 && (hi_lane || lo_lane)) {
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(gr.g_w_col1 + c * HID + jj, hi_lane ? v[c] + v[8 + c] : v[c]);
        }
        if constexpr (HID == 64) {                                // dW1[0][i] = dsdf^T H1
            tmem_ld16(tlane + A::t_w1, v);
            if (gr.g_w_sdf1 && (hi_lane || lo_lane)) atomicAdd(gr.g_w_sdf1 + jj, hi_lane ? v[0] + v[8] : v[0]);
        }
    }
    // M21[j][i] = sum over the samples of dH2[j] H1[i] -> scratch, in a fixed order (the hi lane of a unit writes hi.hi + hi.lo, its lo
    // lane adds lo.hi after a barrier): no atomics, the result does not depend on thread timing
    if (h == 0 && hi_lane) {                                      // warp-uniform: HID is a multiple of 32
#pragma unroll 1
        for (int q = 0; q < HID / 32; ++q) {
            float u[32], v[32];
            tmem_ld32(tlane + A::t_m + 32 * q, u);
            tmem_ld32(tlane + A::t_m + HID + 32 * q, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) scratch[jj * HID + 32 * q + i] = u[i] + v[i];
        }
    }
    grp_sync2(g);
    if (h == 0 && lo_lane) {
#pragma unroll 1
        for (int q = 0; q < HID / 32; ++q) {
            float u[32];
            tmem_ld32(tlane + A::t_m + 32 * q, u);
#pragma unroll
            for (int i = 0; i < 32; ++i) scratch[jj * HID + 32 * q + i] += u[i];
        }
    }
    if constexpr (HID == 32) {                                    // lanes 64 / 72 of t_m = the hi / lo part of dsdf: dW1[0][i] = dsdf^T H1
        if (h == 0 && m >= 64 && m < 96) {
            float u[32], v[32];
            tmem_ld32(tlane + A::t_m, u);
            tmem_ld32(tlane + A::t_m + 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float lo_part = __shfl_sync(0xffffffffu, u[i], 8);
                if (m == 64 && gr.g_w_sdf1) atomicAdd(gr.g_w_sdf1 + i, (u[i] + v[i]) + lo_part);
            }
        }
    }
    grp_sync2(g);
    for (int o = h * 128 + m; o < 2 * kGeo * HID; o += 256) {
        float acc = 0.f;
        if (o < kGeo * HID) {                                     // dW1[1 + gg][i] = sum_j W2[j][48 + gg] M21[j][i]
            const int gg = o / HID, i = o - gg * HID;
            for (int j = 0; j < HID; ++j) acc = fmaf(__ldg(wt.w_col0 + j * kIn2 + kBlob + gg), scratch[j * HID + i], acc);
            if (gr.g_w_sdf1) atomicAdd(gr.g_w_sdf1 + (1 + gg) * HID + i, acc);
        } else {                                                  // dW2[j][48 + gg] = sum_i M21[j][i] W1[1 + gg][i]
            const int o2 = o - kGeo * HID, j = o2 / kGeo, gg = o2 - j * kGeo;
            for (int i = 0; i < HID; ++i) acc = fmaf(scratch[j * HID + i], __ldg(wt.w_sdf1 + (1 + gg) * HID + i), acc);
            if (gr.g_w_col0) atomicAdd(gr.g_w_col0 + j * kIn2 + kBlob + gg, acc);
        }
    }
    grp_sync2(g);                                                 // scratch is operand memory again
}

// sum_k d blob_k / dx * g_k for the 16 bins of one coordinate (Appendix B7 derivative: d out_k / dx = pdf_k - pdf_{k+1} at the
// bin boundaries).  As in stage_oneblob, inside [-0.9, 1.9] only the two boundaries next to x have a non-zero kernel
// value, so three bins carry gradient; elsewhere the general form.
__device__ __forceinline__ float oneblob_dot_grad(float x, const float (&g)[16]) {
    float acc = 0.f;
    if (x > -0.9f && x < 1.9f) {
        float xw = x - floorf(x);
        if (xw >= 1.0f) xw = 0.0f;
        const float t = xw * (float)kNB;
        const int b = (int)t;
        const float ub = (float)b - t;
        const float pb = quartic_pdf(ub / (float)kNB, (float)kNB), pb1 = quartic_pdf((ub + 1.0f) / (float)kNB, (float)kNB);
        const int km = (b - 1) & (kNB - 1), kp = (b + 1) & (kNB - 1);
#pragma unroll
        for (int k = 0; k < kNB; ++k) {
            const float w = (k == km ? -pb : 0.f) + (k == b ? pb - pb1 : 0.f) + (k == kp ? pb1 : 0.f);
            acc = fmaf(w, g[k], acc);
        }
    } else {
        float gb[kNB];
        oneblob_coord_grad<kNB>(x, gb);
#pragma unroll
        for (int k = 0; k < kNB; ++k) acc = fmaf(gb[k], g[k], acc);
    }
    return acc;
}

// BA: additionally writes, per sample, the gradient w.r.t. the GBV texel (dgb [P][4]) and the OneBlob part of the
// gradient w.r.t. the normalised position (dxb [3][P]) for raygrad_walk_kernel: one more tensor-core phase per tile.
// One flag per 128-row tile: does any of its rows (plane index q = s * N + r) lie below its ray's n_live?  A warp per tile.
__global__ void __launch_bounds__(256) tile_live_kernel(long long n_rays, long long P, const int* __restrict__ n_live, unsigned char* __restrict__ flags) {
    const long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (tile >= ws_tiles(P)) return;
    const int lane = threadIdx.x & 31;
    bool any = false;
    const long long q0 = tile * kTile, s0 = q0 / n_rays;
    long long r = q0 - s0 * n_rays + lane, s = s0;
#pragma unroll
    for (int j = 0; j < 4; ++j, r += 32) {
        while (r >= n_rays) { r -= n_rays; ++s; }
        if (q0 + lane + 32 * j < P) any |= s < (long long)__ldg(n_live + r);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[tile] = any ? 1 : 0;
}

// RF_MLP_TRACE (debug builds only: RF_NVCC_DEFS=-DRF_MLP_TRACE): the issuing thread of every group accumulates the clock cycles between
// the marked points of its tile loop; rf_debug_mlp_trace() reads the sums (slot 31 = processed tiles).
#ifdef RF_MLP_TRACE
__device__ unsigned long long g_mlp_trace[32];
#define TR(i) do { if (issuer) { const long long c_ = clock64(); tr_acc[i] += (unsigned long long)(c_ - tr_prev); tr_prev = c_; } } while (0)
#else
#define TR(i) do { } while (0)
#endif

template <int HID, int G, bool BA, bool TMAH>
__global__ void __launch_bounds__(G * 256, 1) mlp_bwd_tc2_kernel(RayK k, Weights wts, const float* __restrict__ feat, long long P,
                                                                 const float* __restrict__ d_raw_tot, const unsigned char* __restrict__ tile_flags,
                                                                 float* __restrict__ dfeat, Grads gr, float* __restrict__ dgb, float* __restrict__ dxb) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bars[G], tma_bars[G];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float w1s[HID];                                                  // W1[0][:], the sdf row (fp32)
    using W = WL<HID>; using A = BwdL<HID>;
    constexpr int HC = HID / 8, NH = HID / 2;
    constexpr uint32_t TCOLS = G * A::tcols;
    // T_A: working accumulator; T_S: [hi HID/2 | lo HID/2] of the narrow operand the next GEMM reads K-major (H1, dRGB, dH2, dH1 in turn)
    constexpr uint32_t T_A = A::t_a, T_S = A::t_s;
    static_assert(T_S + HID <= A::tcols, "operand slot exceeds the group's TMEM columns");
    const int tid = threadIdx.x, m = tid & 127, h = (tid >> 7) & 1, warp = tid >> 5;
    // Warp-uniform copies (shuffle broadcasts, the compiler's uniformity analysis understands them): everything an MMA operand is
    // derived from — group index, shared-memory and TMEM bases, tile liveness — must be provably uniform, otherwise every
    // tcgen05.mma is wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop over the active lanes (~80 cycles per MMA, measured
    // with RF_MLP_TRACE: 41 % of a tile's time was MMA issue)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int g = warp_u >> 3;
    // weights sit after the group regions: the M = 128 MN-major reads of the narrow H / D operands run past their own
    // chunks (rows of the accumulator that are never read back) and must stay inside the allocation
    unsigned char* wsm = smem + G * A::bytes;
    unsigned char* act = smem + g * A::bytes;
    if (warp == 0) tmem_alloc(&tmem_base_s, TCOLS);
    if (tid == 0) { for (int i = 0; i < G; ++i) { mbar_init(&bars[i], 1); mbar_init(&tma_bars[i], 1); } fence_mbar_init(); }
    load_weights<HID>(wsm, wts, G * 256);
    load_w21<HID>(wsm, w1s, wts, G * 256);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0) + (uint32_t)g * A::tcols;
    const uint32_t tlane = tb + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t ts_hi = tlane + T_S, ts_lo = tlane + T_S + NH;
    const bool iwarp = (warp_u & 7) == 0;                                                     // the group's MMA-issuing warp (uniform)
    const bool issuer = iwarp && (tid & 31) == 0;                                             // bulk copies, trace
    uint64_t* bar = &bars[g];
    uint32_t phase = 0;
    unsigned char *x_hi = act + A::c_x_hi * kChunkB, *x_lo = act + A::c_x_lo * kChunkB;
    unsigned char *h1_hi = act + A::c_h1_hi * kChunkB, *h1_lo = act + A::c_h1_lo * kChunkB;
    unsigned char *h2_hi = act + A::c_h2_hi * kChunkB, *h2_lo = act + A::c_h2_lo * kChunkB;
    unsigned char *d_hi = act + A::c_d_hi * kChunkB, *d_lo = act + A::c_d_lo * kChunkB;
    unsigned char *blob_hi = x_hi + kXBlob * kChunkB, *blob_lo = x_lo + kXBlob * kChunkB;
    unsigned char *tail_hi = x_hi + kXTail * kChunkB, *tail_lo = x_lo + kXTail * kChunkB;
    const uint32_t w0h = smem_u32(wsm + W::o_w0h), w0l = smem_u32(wsm + W::o_w0l);
    const uint32_t w2h = smem_u32(wsm + W::o_w2h), w2l = smem_u32(wsm + W::o_w2l), w3h = smem_u32(wsm + W::o_w3h), w3l = smem_u32(wsm + W::o_w3l);
    const uint32_t w21h = smem_u32(wsm + W::o_w21h), w21l = smem_u32(wsm + W::o_w21l);
    const uint32_t xh = smem_u32(x_hi), xl = smem_u32(x_lo), h1h = smem_u32(h1_hi), h1l = smem_u32(h1_lo), h2h = smem_u32(h2_hi), h2l = smem_u32(h2_lo);
    const uint32_t dh = smem_u32(d_hi);
    constexpr uint32_t idH = idesc_bf16(HID, false, false);
    constexpr uint32_t idH_bm = idesc_bf16(HID, false, true), id32_bm = idesc_bf16(32, false, true);
    constexpr uint32_t id16_mm = idesc_bf16(16, true, true), id2H_mm = idesc_bf16(2 * HID, true, true);
    uint32_t wacc = 0;
    int since_flush = 0;
#ifdef RF_MLP_TRACE
    __shared__ unsigned long long tr_s[G][32];
    unsigned long long* tr_acc = tr_s[g];
    if ((tid & 255) < 32) tr_acc[tid & 255] = 0;
    __syncthreads();
    long long tr_prev = clock64();
#endif

    uint64_t* tma_bar = &tma_bars[g];
    uint32_t tma_phase = 0;
    const long long tstep = (long long)gridDim.x * G, n_tiles = ws_tiles(P);
    TilePos tp; tp.init(k.n_rays, k.S, ((long long)blockIdx.x * G + g) * kTile, tstep * kTile);
    // A tile none of whose 128 rows carries an upstream gradient (rows beyond their ray's n_live, composite_bwd_kernel) contributes
    // nothing and is skipped as a whole: tile_live_kernel has reduced n_live to one flag per tile, so the decision is one
    // uniform byte load, issued two tiles ahead (no per-row loads, no vote barrier)
    auto flag_of = [&](long long tile_) -> uint32_t {
        if (tile_ >= n_tiles) return 0u;
        return tile_flags ? (uint32_t)__ldg(tile_flags + tile_) : 1u;
    };
    // hash chunks of tile `tile_` -> chunks 0..3 of X hi / X lo: two bulk copies of 8 KB, completion on tma_bar
    auto fetch_hash = [&](long long tile_) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(feat) + tile_ * kHopTileBytes;
        mbar_expect_tx(tma_bar, kHopTileBytes);
        bulk_g2s(x_hi, src, kHopTileBytes / 2, tma_bar);
        bulk_g2s(x_lo, src + kHopTileBytes / 2, kHopTileBytes / 2, tma_bar);
    };
    TileHalf t;
    float4 dr_next = make_float4(0.f, 0.f, 0.f, 0.f);
    long long tile = (long long)blockIdx.x * G + g;
    auto load_inputs = [&](const TilePos& pos, long long tile_) {                             // inputs of an alive tile
        load_half<BA, TMAH>(t, feat, P, tile_, m, h);
        dr_next = (tile_ * kTile + m < P) ? __ldg(reinterpret_cast<const float4*>(d_raw_tot) + pos.raw_index(m)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // the next alive tile has its inputs loaded / its hash chunks fetched while the current one is computed, the one after it is
    // pulled towards L2; flags are made warp-uniform for the compiler (see above)
    TilePos tn = tp; tn.next();
    bool alive = __shfl_sync(0xffffffffu, flag_of(tile), 0) != 0;
    bool alive_n = __shfl_sync(0xffffffffu, flag_of(tile + tstep), 0) != 0;
    if (alive) { if (TMAH && issuer) fetch_hash(tile); load_inputs(tp, tile); }
    if (alive_n && h == 0) prefetch_tile(feat, P, tile + tstep, m, true);
    TilePos tnn = tn; tnn.next();
    uint32_t fl_nn = flag_of(tile + 2 * tstep);                                               // consumed one iteration later
    for (; tile < n_tiles; tile += tstep, tp = tn, tn = tnn, tnn.next()) {
        const long long q = tile * kTile + m;                                                 // plane index s * N + r
        const bool live = q < P;
        const long long tile_n = tile + tstep, tile_nn = tile_n + tstep;
        const bool alive_nn = __shfl_sync(0xffffffffu, fl_nn, 0) != 0;
        fl_nn = flag_of(tile_nn + tstep);
        if (alive_nn && h == 0) prefetch_tile(feat, P, tile_nn, m, true);
        const bool alive_cur = alive, alive_nx = alive_n;
        alive = alive_n; alive_n = alive_nn;                                                  // shifted for the next iteration
        if (!alive_cur) {                                                                     // nothing of this tile is needed
            if (alive_nx) { if (TMAH && issuer) fetch_hash(tile_n); load_inputs(tn, tile_n); }
            TR(24);
            continue;
        }
        TR(0);
        const float4 dr = dr_next;
        const float xg1 = h ? t.xb : t.x0, xg2 = h ? t.g.x : t.xb;                            // BA: see load_half
        {                                                                                     // X row m, half h
            if (!TMAH) {                                                                      // ready operand chunks 2h, 2h + 1
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t off = chunk_off(128, m, 2 * h + c);
                    *reinterpret_cast<uint4*>(x_hi + off) = t.hh[c];
                    *reinterpret_cast<uint4*>(x_lo + off) = t.hl[c];
                }
            }
            if (h == 0) {
                if (live) { stage_oneblob(t.x0, blob_hi, blob_lo, m, 0); stage_oneblob(t.x1, blob_hi, blob_lo, m, 2); }
                else { for (int c = 0; c < 4; ++c) stage_zero(blob_hi, blob_lo, m, c); }
            } else {
                if (live) stage_oneblob(t.x0, blob_hi, blob_lo, m, 4);
                else { stage_zero(blob_hi, blob_lo, m, 4); stage_zero(blob_hi, blob_lo, m, 5); }
                float t_add, cin, d0, d1;
                tsdf_terms(k, 0, t.g.x, t_add, cin, d0, d1);
                float v1[8] = {t.g.y, t.g.z, t.g.w, cin, 0.f, 0.f, 0.f, 0.f};                   // tail: gbv rgb | decoder tsdf input (the geo features: W21)
                stage8(tail_hi, tail_lo, m, 0, v1);
                stage_zero(tail_hi, tail_lo, m, 1);
            }
        }
        if (alive_nx) load_inputs(tn, tile_n);                                                // next tile's inputs
        TR(1);
        fence_async_smem(); fence_before_sync(); grp_sync2(g);
        TR(2);
        if (TMAH && iwarp) {                                                                  // this tile's hash chunks have landed
            if (!mbar_wait(tma_bar, tma_phase)) __trap();
        }
        if (TMAH) tma_phase ^= 1u;
        TR(3);
        if (iwarp && elect_one()) {                                                           // 1: H1 = X1 W0^T
            fence_after_sync();
            uint32_t acc = 0;
            mma_kk<kXCh / 2>(tb + T_A, xh, xl, w0h, w0l, HID, idH, acc);
            commit(bar);
        }
        TR(4);
        grp_wait(bar, phase);
        TR(5);
        const uint32_t mask1 = relu_half<NH>(tlane + T_A + NH * h, h1_hi, h1_lo, m, (HC / 2) * h, ts_hi, ts_lo);
        TR(6);
        tmem_st_wait(); fence_async_smem(); fence_before_sync(); grp_sync2(g);
        TR(7);
        if (iwarp && elect_one()) {                                                                         // 2: H2 = X2' W2^T + H1 W21^T
            fence_after_sync();
            uint32_t acc = 0;
            mma_kk<4>(tb + T_A, xh + kXBlob * kChunkB, xl + kXBlob * kChunkB, w2h, w2l, HID, idH, acc);
            mma_ts<HC / 2>(tb + T_A, tb + T_S, tb + T_S + NH, w21h, w21l, HID, idH, acc);
            commit(bar);
        }
        TR(8);
        grp_wait(bar, phase);
        TR(9);
        uint32_t mask2;
        {                                                                                     // H2 -> shared memory only (no GEMM reads it K-major)
            uint32_t mk = 0;
#pragma unroll
            for (int qq = 0; qq < NH / 16; ++qq) {
                float v[16];
                tmem_ld16(tlane + T_A + NH * h + 16 * qq, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) { if (v[i] > 0.f) mk |= (1u << (16 * qq + i)); v[i] = fmaxf(v[i], 0.f); }
                stage8(h2_hi, h2_lo, m, (HC / 2) * h + 2 * qq, v); stage8(h2_hi, h2_lo, m, (HC / 2) * h + 2 * qq + 1, v + 8);
            }
            mask2 = mk;
        }
        if (h) {                                                                              // dRGB (upstream of :344)
            float v0[8] = {dr.x, dr.y, dr.z, 0.f, 0.f, 0.f, 0.f, 0.f};
            stage8_both(d_hi, d_lo, m, 0, v0, ts_hi, ts_lo);
        } else {                                                                              // K = 16: the upper 8 columns of the operand
            tmem_st4(ts_hi + 4, make_uint4(0, 0, 0, 0)); tmem_st4(ts_lo + 4, make_uint4(0, 0, 0, 0));
        }
        TR(10);
        tmem_st_wait(); fence_async_smem(); fence_before_sync(); grp_sync2(g);
        TR(11);
        if (iwarp && elect_one()) {                                                                         // 3
            fence_after_sync();
            uint32_t acc = 0;
            mma_ts_m<1>(tb + T_A, tb + T_S, tb + T_S + NH, w3h, w3l, 16, idH_bm, acc);        // dH2pre = dRGB W3
            uint32_t a3 = wacc;
            mma_mm1(tb + A::t_w3, h2h, dh, id16_mm, a3);                                      // dW3^T += H2^T dRGB
            commit(bar);
        }
        TR(12);
        grp_wait(bar, phase);
        TR(13);
        masked_half<NH>(tlane + T_A + NH * h, h2_hi, h2_lo, m, (HC / 2) * h, mask2, ts_hi, ts_lo);   // dH2 over H2
        if (h) {                                                                              // D <- dsdf (its MMAs have completed)
            float v0[8] = {dr.w, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            stage8(d_hi, d_lo, m, 0, v0);
        }
        TR(14);
        tmem_st_wait(); fence_async_smem(); fence_before_sync(); grp_sync2(g);
        TR(15);
        if (iwarp && elect_one()) {                                                                         // 4
            fence_after_sync();
            uint32_t acc = 0;
            mma_ts_m<HC / 2>(tb + T_A, tb + T_S, tb + T_S + NH, w21h, w21l, HID, idH_bm, acc);   // dH1pre = dH2 W21 (+ dsdf W1[0] in the epilogue)
            uint32_t am = wacc;
            mma_mm1(tb + A::t_m, h2h, h1h, id2H_mm, am);                                      // M21 += dH2^T H1 (hidden 32: rows 64.. = dsdf^T H1)
            if constexpr (HID == 64) {
                uint32_t a1 = wacc;
                mma_mm1(tb + A::t_w1, h1h, dh, id16_mm, a1);                                  // dW1[0] += dsdf^T H1
                uint32_t a2 = wacc;
                mma_wx(tb + A::t_w2, h2h, xh + kXBlob * kChunkB, xl + kXBlob * kChunkB, idesc_bf16(kK2, true, true), a2);   // dW2 += dH2^T X2'
            }
            commit(bar);
        }
        TR(16);
        grp_wait(bar, phase);
        TR(17);
        masked_half<NH>(tlane + T_A + NH * h, h1_hi, h1_lo, m, (HC / 2) * h, mask1, ts_hi, ts_lo, dr.w, w1s + NH * h);   // dH1 over H1
        TR(18);
        tmem_st_wait(); fence_async_smem(); fence_before_sync(); grp_sync2(g);
        TR(19);
        if (iwarp && elect_one()) {                                                                         // 5
            fence_after_sync();
            uint32_t acc = 0;
            if constexpr (BA) {
                // d [hash | blob(x, y)] = dH1 W0[:, 0..63] (+ dH2 W2[:, blob(x, y)]): 64 columns from T_A, over the operand slot
                constexpr uint32_t id64_bm = idesc_bf16(64, false, true);
                mma_km<HC / 2>(tb + T_A, h1h, h1l, w0h, w0l, HID, id64_bm, acc);
                uint32_t one = 1;
                mma_km<HC / 2>(tb + T_A + 32, h2h, h2l, w2h, w2l, HID, id32_bm, one);
            } else {
                mma_ts_m<HC / 2>(tb + T_A, tb + T_S, tb + T_S + NH, w0h, w0l, HID, id32_bm, acc); // d hash = dH1 W0[:, 0..31]
            }
            uint32_t a0 = wacc;
            mma_wx(tb + A::t_w0, h1h, xh, xl, idesc_bf16(kKX, true, true), a0);               // dW0 += dH1^T X1 (hidden 32: and dW2 += dH2^T X)
            commit(bar);
        }
        wacc = 1;
        TR(20);
        grp_wait(bar, phase);
        TR(21);
        // every MMA that reads this tile's X operand has completed: the next alive tile's hash chunks may land on it
        if (TMAH && issuer && alive_nx) fetch_hash(tile_n);
        {
            float dx[16];
            tmem_ld16(tlane + T_A + 16 * h, dx);
            if (live) {                                                                       // dfeat [4 quads][P][8]: quads 2h, 2h + 1
                float4* dj = reinterpret_cast<float4*>(dfeat) + ((long long)(2 * h) * P + q) * 2;
                dj[0] = make_float4(dx[0], dx[1], dx[2], dx[3]); dj[1] = make_float4(dx[4], dx[5], dx[6], dx[7]);
                dj[2 * P] = make_float4(dx[8], dx[9], dx[10], dx[11]); dj[2 * P + 1] = make_float4(dx[12], dx[13], dx[14], dx[15]);
            }
        }
        if constexpr (BA) {
            {                                                                                 // OneBlob gradient of x (h = 0) / y (h = 1)
                float gb[16];
                tmem_ld16(tlane + T_A + 32 + 16 * h, gb);
                if (live) dxb[(long long)h * P + q] = oneblob_dot_grad(xg1, gb);
            }
            fence_before_sync(); grp_sync2(g);
            if (iwarp && elect_one()) {
                // d [blob(z) | tail] -> columns 0..31, from both nets (the geo features are not inputs here: their part of the
                // gradient went through W21)
                fence_after_sync();
                uint32_t acc = 0;
                mma_km<HC / 2>(tb + T_A, h1h, h1l, w0h + 8 * HID * 16, w0l + 8 * HID * 16, HID, id32_bm, acc);
                mma_km<HC / 2>(tb + T_A, h2h, h2l, w2h + 4 * HID * 16, w2l + 4 * HID * 16, HID, id32_bm, acc);
                commit(bar);
            }
            grp_wait(bar, phase);
            if (h == 0) {
                float gb[16];
                tmem_ld16(tlane + T_A, gb);
                if (live) dxb[2 * P + q] = oneblob_dot_grad(xg2, gb);
            } else {
                float ta[16];
                tmem_ld16(tlane + T_A + 16, ta);                                              // d tail: [gbv r, g, b | decoder tsdf input | 0]
                float t_add, cin, dg_add, dg_cin;
                tsdf_terms(k, 0, xg2, t_add, cin, dg_add, dg_cin);
                // GBV texel gradient: colour-net inputs + the residual adds (:344-345); tsdf through the decoder input and the add
                if (live) reinterpret_cast<float4*>(dgb)[q] = make_float4(ta[3] * dg_cin + dr.w * dg_add, ta[0] + dr.x, ta[1] + dr.y, ta[2] + dr.z);
            }
        }
        TR(22);
#ifdef RF_MLP_TRACE
        if (issuer) tr_acc[31] += 1;
#endif
        if (++since_flush == kFlushTiles) {
            fence_before_sync();
            flush_wgrads3<HID>(tlane, gr, wts, reinterpret_cast<float*>(h2_hi), g, m, h); wacc = 0; since_flush = 0;
        }
        fence_before_sync();
    }
    if (wacc) flush_wgrads3<HID>(tlane, gr, wts, reinterpret_cast<float*>(h2_hi), g, m, h);
    TR(25);
#ifdef RF_MLP_TRACE
    if (issuer) for (int i = 0; i < 32; ++i) atomicAdd(&g_mlp_trace[i], tr_acc[i]);
#endif
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, TCOLS);
}

template <int HID, int G> static size_t fwd_bytes() { return (size_t)G * FwdL<HID>::bytes + WL<HID>::total_bwd; }
template <int HID, int G> static size_t bwd_bytes() { return (size_t)G * BwdL<HID>::bytes + WL<HID>::total_bwd; }

template <int HID, int G>
static int launch_fwd_g(const RayK& k, const Weights& w, const float* feat, long long P, int variant, float* raw, cudaStream_t s) {
    auto fn = mlp_fwd_tc_kernel<HID, G>;
    size_t sm = fwd_bytes<HID, G>();
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(mlp_fwd_tc, %zu B): %s", sm, cudaGetErrorString(e));
    long long tiles = (P + kTile - 1) / kTile;
    int blocks = (int)std::min<long long>((tiles + G - 1) / G, (long long)num_sms());
    ProfScope ps(RF_PROF_MLP_FWD, s);
    fn<<<blocks, G * 128, sm, s>>>(k, w, feat, P, variant, raw);
    RF_CHECK_LAUNCH("mlp_fwd_tc_kernel");
    return 0;
}
template <int HID, int G>
static int launch_bwd_g(const RayK& k, const Weights& w, const float* feat, long long P, const float* d_raw_tot, const int* n_live,
                        unsigned char* tile_flags, float* dfeat, const Grads& gr, float* dgb, float* dxb, cudaStream_t s) {
    const bool ba = dgb != nullptr;
    // RF_BWD_HASH_TMA=0 moves the hash operand chunks through registers (LDG.128 -> STS.128) instead of TMA bulk copies
    static const bool tma = [] { const char* e = getenv("RF_BWD_HASH_TMA"); return e ? atoi(e) != 0 : true; }();
    auto fn = ba ? (tma ? mlp_bwd_tc2_kernel<HID, G, true, true> : mlp_bwd_tc2_kernel<HID, G, true, false>)
                 : (tma ? mlp_bwd_tc2_kernel<HID, G, false, true> : mlp_bwd_tc2_kernel<HID, G, false, false>);
    size_t sm = bwd_bytes<HID, G>();
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return set_error((int)e, "cudaFuncSetAttribute(mlp_bwd_tc2, %zu B): %s", sm, cudaGetErrorString(e));
    long long tiles = (P + kTile - 1) / kTile;
    int blocks = (int)std::min<long long>((tiles + G - 1) / G, (long long)num_sms());
    ProfScope ps(RF_PROF_MLP_BWD, s);
    if (n_live) {
        tile_live_kernel<<<(unsigned)((tiles * 32 + 255) / 256), 256, 0, s>>>(k.n_rays, P, n_live, tile_flags);
        RF_CHECK_LAUNCH("tile_live_kernel");
    }
    fn<<<blocks, G * 256, sm, s>>>(k, w, feat, P, d_raw_tot, n_live ? tile_flags : nullptr, dfeat, gr, dgb, dxb);
    RF_CHECK_LAUNCH("mlp_bwd_tc2_kernel");
    return 0;
}

}  // namespace

#ifdef RF_MLP_TRACE
extern "C" int rf_debug_mlp_trace(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_mlp_trace, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_mlp_trace, z, sizeof(z)); }
    return 0;
}
#endif

int launch_encode(const RayK& k, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o, const float* rays_d,
                  const float* z_vals, long long P, float* feat, cudaStream_t s);
struct RayGradArgs;
int launch_scatter(const RayK& k, const GridDev& hg, long long P, const float* feat, const float* dfeat, const int* n_live, float* g_hash,
                   float* g_rep, const RayGradArgs* rg, cudaStream_t s);
int launch_encode_points(const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* x, long long n, float* feat, cudaStream_t s);

bool tc_supported(const RayK& k, int hidden) { return k.n_hash_out == 32 && (hidden == 32 || hidden == 64); }

// workspace `feat`: ws_floats(P) floats (ray_common.cuh) written here and read back by launch_bwd_tc
int launch_fwd_tc(const RayK& k, int hidden, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* rays_o,
                  const float* rays_d, const float* z_vals, long long P, float* raw, float* feat, cudaStream_t s) {
    int rc = launch_encode(k, hg, gg, p, rays_o, rays_d, z_vals, P, feat, s);
    if (rc) return rc;
    Weights w{p->w_sdf0, p->w_sdf1, p->w_col0, p->w_col1};
    // groups per SM: hidden 64 -> 3 (TMEM 3 x 160 columns).  Hidden 32 would fit 5 (5 x 96 columns) but 640 threads cap the kernel at 96
    // registers: measured 3.46 ms with 5 groups against 3.32 ms with 4 (bench config 2)
    return hidden == 64 ? launch_fwd_g<64, 3>(k, w, feat, P, 0, raw, s) : launch_fwd_g<32, 4>(k, w, feat, P, 0, raw, s);
}

// point queries through the same kernels: n points = n rays of one sample; variant selects the tsdf handling
int launch_points_tc(RayK k, int hidden, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, const float* x, long long n,
                     int variant, float* raw, float* feat, cudaStream_t s) {
    k.n_rays = n; k.S = 1;
    int rc = launch_encode_points(hg, gg, p, x, n, feat, s);
    if (rc) return rc;
    Weights w{p->w_sdf0, p->w_sdf1, p->w_col0, p->w_col1};
    return hidden == 64 ? launch_fwd_g<64, 3>(k, w, feat, n, variant, raw, s) : launch_fwd_g<32, 4>(k, w, feat, n, variant, raw, s);
}

// dfeat: 2L * P floats of scratch ([4 quads][P][8]), then (ray gradients only) 4P + 3P floats for the GBV-texel and OneBlob
// gradients, then scatter_scratch_floats() floats for the table replicas, then one byte per tile (liveness flags)
int launch_bwd_tc(const RayK& k, int hidden, const GridDev& hg, const GridDev& gg, const rf_ray_params* p, long long P, const float* feat,
                  const float* d_raw_tot, const int* n_live, float* dfeat, const Grads& gr, float* g_rays_o, float* g_rays_d, cudaStream_t s) {
    Weights w{p->w_sdf0, p->w_sdf1, p->w_col0, p->w_col1};
    const bool ba = g_rays_o || g_rays_d;
    float* dgb = ba ? dfeat + 2ll * hg.n_levels * P : nullptr;
    float* dxb = ba ? dgb + 4 * P : nullptr;
    float* rep = ba ? dxb + ((3 * P + 3) & ~3ll) : dfeat + 2ll * hg.n_levels * P;     // replicas are float2 / float4 accessed
    unsigned char* tile_flags = reinterpret_cast<unsigned char*>(rep + scatter_scratch_floats(hg, k.n_rays));      // ws_tiles(P) bytes
    int rc = hidden == 64 ? launch_bwd_g<64, 1>(k, w, feat, P, d_raw_tot, n_live, tile_flags, dfeat, gr, dgb, dxb, s)
                          : launch_bwd_g<32, 2>(k, w, feat, P, d_raw_tot, n_live, tile_flags, dfeat, gr, dgb, dxb, s);
    if (rc) return rc;
    if (ba) return launch_scatter_raygrad(k, hg, gg, p, P, feat, dfeat, dgb, dxb, n_live, gr.g_hash, rep, g_rays_o, g_rays_d, s);
    if (gr.g_hash) rc = launch_scatter(k, hg, P, feat, dfeat, n_live, gr.g_hash, rep, nullptr, s);
    return rc;
}

}  // namespace rf
