// grid_encode.cuh — device-side multiresolution grid (hash / dense) and OneBlob encoders.
//
// Semantics follow tiny-cuda-nn (the reference's un-vendored dependency, model/encodings.py:39-50,67-74 and
// model/scene_rep.py:60-93): SURVEY.md Appendix B1-B7.  Index arithmetic is 32-bit two's-complement exact
// (uint32 wrap of the cell, wrap-around multiply / xor, modulo the level size) so that samples outside the unit
// cube address the same entries tiny-cuda-nn would (Appendix A18).
#pragma once
#include "rf_common.cuh"

namespace rf {

struct GridDev {               // kernel-side copy of rf_grid_desc
    int   n_levels, n_features, is_hash;
    unsigned hash_pow2_mask;   // bit l: level l uses the prime hash and its table size is a power of two
    unsigned dense_mask;       // bit l: level l is indexed densely (size >= res^3)
    float scale[RF_MAX_LEVELS];
    unsigned res[RF_MAX_LEVELS];
    unsigned size[RF_MAX_LEVELS];
    unsigned offset[RF_MAX_LEVELS + 1];
};

static inline GridDev to_dev(const rf_grid_desc* d) {
    GridDev g;
    g.n_levels = d->n_levels; g.n_features = d->n_features; g.is_hash = d->is_hash;
    for (int i = 0; i < RF_MAX_LEVELS; ++i) { g.scale[i] = d->scale[i]; g.res[i] = d->resolution[i]; g.size[i] = d->size[i]; }
    for (int i = 0; i <= RF_MAX_LEVELS; ++i) g.offset[i] = d->offset[i];
    g.hash_pow2_mask = 0; g.dense_mask = 0;
    for (int l = 0; l < d->n_levels && l < RF_MAX_LEVELS; ++l) {          // the rule of grid_index / CornerIndexer::init
        const unsigned long long size = d->size[l], res = d->resolution[l];
        unsigned long long stride = 1;
        for (int k = 0; k < 3; ++k) if (stride <= size) stride *= res;
        const bool hashed = d->is_hash && size < stride;
        if (hashed && size && (size & (size - 1)) == 0) g.hash_pow2_mask |= 1u << l;
        if (!hashed && res * res * res <= size) g.dense_mask |= 1u << l;
    }
    return g;
}

// Appendix B2: pos = fmaf(scale, x, 0.5); cell = (uint32)(int)floor(pos); frac = pos - floor(pos)
__device__ __forceinline__ void pos_fract(float x, float scale, unsigned& cell, float& frac) {
    float pos = fmaf(scale, x, 0.5f);
    float fl = floorf(pos);
    cell = (unsigned)(int)fl;
    frac = pos - fl;
}

// Appendix B3: index of one corner within a level.
__device__ __forceinline__ unsigned grid_index(bool is_hash, unsigned size, unsigned res, unsigned cx, unsigned cy, unsigned cz) {
    const unsigned c[3] = {cx, cy, cz};
    unsigned stride = 1, index = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (stride <= size) { index += c[d] * stride; stride *= res; }
    }
    if (is_hash && size < stride) index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    return index % size;
}

// Same index, without the generic 32-bit modulo on the common paths: power-of-two sizes (every hashed level: T = 2^k)
// reduce with a mask, and a dense index only needs the modulo when the sample lies outside the grid (A18).
__device__ __forceinline__ unsigned grid_index_fast(bool is_hash, unsigned size, unsigned res, unsigned cx, unsigned cy, unsigned cz) {
    const unsigned c[3] = {cx, cy, cz};
    unsigned stride = 1, index = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (stride <= size) { index += c[d] * stride; stride *= res; }
    }
    if (is_hash && size < stride) index = cx ^ (cy * 2654435761u) ^ (cz * 805459861u);
    if ((size & (size - 1u)) == 0u) return index & (size - 1u);
    return (index < size) ? index : index % size;
}

// The 8 corner indices of one cell at once (same arithmetic as grid_index, Appendix B3, shared between the corners):
// hashed levels reuse cy*P1, cz*P2 (and (c+1)*P = c*P + P in uint32), dense levels add constant offsets to one base
// index; a power-of-two size reduces with a mask, any other size only pays the modulo when the index is out of range.
struct CornerIndexer {
    unsigned size, mask, s2, s3;
    bool hashed, pow2;
    __device__ __forceinline__ void init(bool is_hash, unsigned size_, unsigned res) {
        size = size_; s2 = 0; s3 = 0;
        unsigned stride = res;                                   // after dimension 0 (stride 1 <= size always)
        if (stride <= size) { s2 = stride; stride *= res; if (stride <= size) { s3 = stride; stride *= res; } }
        hashed = is_hash && size < stride;
        pow2 = (size & (size - 1u)) == 0u; mask = size - 1u;
    }
    __device__ __forceinline__ void cell(unsigned cx, unsigned cy, unsigned cz, unsigned (&idx)[8]) const {
        if (hashed) {
            const unsigned a[2] = {cx, cx + 1u};
            const unsigned b0 = cy * 2654435761u, c0 = cz * 805459861u;
            const unsigned b[2] = {b0, b0 + 2654435761u}, c[2] = {c0, c0 + 805459861u};
#pragma unroll
            for (int k = 0; k < 8; ++k) idx[k] = a[k & 1] ^ b[(k >> 1) & 1] ^ c[(k >> 2) & 1];
        } else {
            const unsigned base = cx + cy * s2 + cz * s3;
#pragma unroll
            for (int k = 0; k < 8; ++k) idx[k] = base + (unsigned)(k & 1) + ((k & 2) ? s2 : 0u) + ((k & 4) ? s3 : 0u);
        }
        if (pow2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) idx[k] &= mask;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) if (idx[k] >= size) idx[k] %= size;
        }
    }
};

// The 8 corner indices with the two common cases specialised (same values as CornerIndexer / grid_index, Appendix B3):
//   * hashed level with a power-of-two table: (a ^ b ^ c) & mask == (a & mask) ^ (b & mask) ^ (c & mask), one 3-input
//     logic op per corner;
//   * dense level, cell strictly inside the grid: base + constant offsets, no wrap-around modulo;
// everything else (samples outside the unit cube on dense levels, odd table sizes) takes the general path.
__device__ __forceinline__ void cell_indices(const GridDev& g, int l, unsigned cx, unsigned cy, unsigned cz, unsigned (&idx)[8]) {
    const unsigned size = g.size[l], res = g.res[l];
    if ((g.hash_pow2_mask >> l) & 1u) {
        const unsigned mask = size - 1u;
        const unsigned a[2] = {cx & mask, (cx + 1u) & mask};
        const unsigned b0 = cy * 2654435761u, c0 = cz * 805459861u;
        const unsigned b[2] = {b0 & mask, (b0 + 2654435761u) & mask}, c[2] = {c0 & mask, (c0 + 805459861u) & mask};
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = a[k & 1] ^ b[(k >> 1) & 1] ^ c[(k >> 2) & 1];
    } else if (((g.dense_mask >> l) & 1u) && cx + 1u < res && cy + 1u < res && cz + 1u < res && cx + 1u != 0u && cy + 1u != 0u && cz + 1u != 0u) {
        const unsigned s2 = res, s3 = res * res, base = cx + cy * s2 + cz * s3;
#pragma unroll
        for (int k = 0; k < 8; ++k) idx[k] = base + (unsigned)(k & 1) + ((k & 2) ? s2 : 0u) + ((k & 4) ? s3 : 0u);
    } else {
        CornerIndexer ci; ci.init(g.is_hash != 0, size, res);
        ci.cell(cx, cy, cz, idx);
    }
}

// Corner weight, Appendix B4 (weight = 1; for dim: weight *= frac or 1-frac).
__device__ __forceinline__ float corner_weight(int corner, float fx, float fy, float fz) {
    float w = 1.0f;
    w *= (corner & 1) ? fx : 1.0f - fx;
    w *= (corner & 2) ? fy : 1.0f - fy;
    w *= (corner & 4) ? fz : 1.0f - fz;
    return w;
}

// ---- OneBlob (Appendix B7) -------------------------------------------------------------------------------
__device__ __forceinline__ float quartic_cdf(float t, float n_bins) {
    float u = t * n_bins;
    float u2 = u * u, u4 = u2 * u2;
    return fminf(fmaxf((15.0f / 16.0f) * u * (1.0f - (2.0f / 3.0f) * u2 + (1.0f / 5.0f) * u4) + 0.5f, 0.0f), 1.0f);
}
// d/dt of the clamped cdf: the quartic kernel (zero where the clamp is active)
__device__ __forceinline__ float quartic_pdf(float t, float n_bins) {
    float u = t * n_bins;
    float m = fmaxf(1.0f - u * u, 0.0f);
    return (15.0f / 16.0f) * m * m * n_bins;
}
__device__ __forceinline__ float blob_left_cdf(float x, int k, int n_bins) {
    float l = (float)k / (float)n_bins - x;
    float nb = (float)n_bins;
    return quartic_cdf(l, nb) + quartic_cdf(l - 1.0f, nb) + quartic_cdf(l + 1.0f, nb);
}
__device__ __forceinline__ float blob_left_pdf(float x, int k, int n_bins) {
    float l = (float)k / (float)n_bins - x;
    float nb = (float)n_bins;
    return quartic_pdf(l, nb) + quartic_pdf(l - 1.0f, nb) + quartic_pdf(l + 1.0f, nb);
}
// All bins of one coordinate: out[k] = L_{k+1} - L_k with L_nbins := L_0 + 1
template <int NB>
__device__ __forceinline__ void oneblob_coord(float x, float (&out)[NB]) {
    float first = blob_left_cdf(x, 0, NB);
    float left = first;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        float right = (k == NB - 1) ? first + 1.0f : blob_left_cdf(x, k + 1, NB);
        out[k] = right - left;
        left = right;
    }
}
// d out[k] / d x = -(pdf_{k+1} - pdf_k)  (pdf of the wrapped boundary equals pdf_0)
template <int NB>
__device__ __forceinline__ void oneblob_coord_grad(float x, float (&dout)[NB]) {
    float first = blob_left_pdf(x, 0, NB);
    float left = first;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        float right = (k == NB - 1) ? first : blob_left_pdf(x, k + 1, NB);
        dout[k] = left - right;
        left = right;
    }
}

}  // namespace rf
