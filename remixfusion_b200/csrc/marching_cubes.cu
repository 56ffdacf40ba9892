// marching_cubes.cu — N4 (second half): iso-surface extraction on the device over the lattice volume that
// remixfusion_b200/lattice.py: query_lattice() leaves in HBM.
//
// Follows the algorithm of the reference's own extractor, thirdparty/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp
// (`mcubes.marching_cubes(volume, isovalue, truncation)`, utils.py:169): a DUAL-grid marching cubes —
//   * a cube is centred on every lattice point; its corners sit at +-0.5 voxel and take the 8-voxel average
//     (trilerp :94-118: all eight weights are 0.5^3, accumulated in the reference's order), valid only if every one of the
//     eight voxels exists and has |d| < truncation (get_voxel :76-92);
//   * case index / vertex-on-edge numbering of extract_isosurface_at_position :139-244, the classic 256-case table, linear
//     interpolation with the reference's three short-cuts (vertexInterp :120-137), the `thresh = 10` consistency checks (:183-202);
//   * triangles are emitted cell by cell in the reference's i, j, k order (prefix sum of per-cell counts), so the triangle
//     soup equals the reference's `results` vector; welding (the reference's merge_close_vertices) happens on the host side
//     of the ABI with device-wide sort / unique on exact keys instead of a tolerance hash.
// Kernels: mc_corner_kernel (coalesced sweep of the volume, one corner per thread, z fastest), mc_cell_kernel<false> (count)
// and mc_cell_kernel<true> (emit).
#include "rf_common.cuh"

namespace rf {
namespace {

// The classic marching-cubes case table (Lorensen & Cline / Bourke): up to five triangles per case, each vertex the
// number 0..11 of a cube edge, packed one nibble per entry (0xf terminates) — one 8-byte constant load per cell.
__constant__ unsigned long long kTriTable[256] = {
    0xffffffffffffffffull, 0xfffffffffffff380ull, 0xfffffffffffff910ull, 0xffffffffff189381ull,
    0xfffffffffffffa21ull, 0xffffffffffa21380ull, 0xffffffffff920a29ull, 0xfffffff89a8a2382ull,
    0xfffffffffffff2b3ull, 0xffffffffff0b82b0ull, 0xffffffffffb32091ull, 0xfffffffb89b912b1ull,
    0xffffffffff3ab1a3ull, 0xfffffffab8a801a0ull, 0xfffffff9ab9b3093ull, 0xffffffffffb8aa89ull,
    0xfffffffffffff874ull, 0xffffffffff437034ull, 0xffffffffff748910ull, 0xfffffff137174914ull,
    0xffffffffff748a21ull, 0xfffffffa21403743ull, 0xfffffff748209a29ull, 0xffff4973727929a2ull,
    0xffffffffff2b3748ull, 0xfffffff40242b74bull, 0xfffffffb32748109ull, 0xffff1292b9b49b74ull,
    0xfffffff487ab31a3ull, 0xffff4b7401b41ab1ull, 0xffff30bab9b09874ull, 0xfffffffab99b4b74ull,
    0xfffffffffffff459ull, 0xffffffffff380459ull, 0xffffffffff051450ull, 0xfffffff513538458ull,
    0xffffffffff459a21ull, 0xfffffff594a21803ull, 0xfffffff204245a25ull, 0xffff8434535235a2ull,
    0xffffffffffb32459ull, 0xfffffff594b802b0ull, 0xfffffffb32510450ull, 0xffff584b82852512ull,
    0xfffffff45931ab3aull, 0xffffab81a8180594ull, 0xffff30bab5b05045ull, 0xfffffffb8aa85845ull,
    0xffffffffff975879ull, 0xfffffff375359039ull, 0xfffffff751710870ull, 0xffffffffff753351ull,
    0xfffffff21a759879ull, 0xffff37503505921aull, 0xffff25a758528208ull, 0xfffffff7533525a2ull,
    0xfffffff2b3987597ull, 0xffffb72029279759ull, 0xffff751871810b32ull, 0xfffffff51771b12bull,
    0xffffb3a31a758859ull, 0xf0aba010b7905075ull, 0xf07570805a30b0abull, 0xffffffffff5b75abull,
    0xfffffffffffff56aull, 0xffffffffff6a5380ull, 0xffffffffff6a5109ull, 0xfffffff6a5891381ull,
    0xffffffffff162561ull, 0xfffffff803621561ull, 0xfffffff620609569ull, 0xffff823625285895ull,
    0xffffffffff56ab32ull, 0xfffffff56a02b80bull, 0xfffffff6a5b32910ull, 0xffffb892b92916a5ull,
    0xfffffff315356b36ull, 0xffff6b51505b0b80ull, 0xffff9505606306b3ull, 0xfffffff89bb96956ull,
    0xffffffffff8746a5ull, 0xfffffffa56374034ull, 0xfffffff7486a5091ull, 0xffff49737179156aull,
    0xfffffff874156216ull, 0xffff743403625521ull, 0xffff620560509748ull, 0xf962695923497937ull,
    0xfffffff56a4872b3ull, 0xffffb720242746a5ull, 0xffff6a5b32874910ull, 0xf6a54b7b492b9129ull,
    0xffff6b51535b3748ull, 0xfb404b7b016b5b15ull, 0xf74836b630560950ull, 0xffff9b7974b96956ull,
    0xffffffffffa4694aull, 0xfffffff380a946a4ull, 0xfffffff04606a10aull, 0xffffa16468618138ull,
    0xfffffff462421941ull, 0xffff462942921803ull, 0xffffffffff624420ull, 0xfffffff624428238ull,
    0xfffffff32b46a94aull, 0xffff6a4a94b82280ull, 0xffffa164606102b3ull, 0xf1b8b12184a16146ull,
    0xffff36b319639469ull, 0xf14641916b0181b8ull, 0xfffffff4600636b3ull, 0xffffffffff86b846ull,
    0xfffffffa98a876a7ull, 0xffffa76a907a0370ull, 0xffff0818717a176aull, 0xfffffff37117a76aull,
    0xffff768981861621ull, 0xf937390976192962ull, 0xfffffff206607087ull, 0xffffffffff276237ull,
    0xffff76898a86ab32ull, 0xf7a9a76790b72702ull, 0xfb32a767a1871081ull, 0xffff17616a71b12bull,
    0xf63136b619768698ull, 0xffffffffff76b190ull, 0xffff06b0b3607087ull, 0xfffffffffffff6b7ull,
    0xfffffffffffffb67ull, 0xffffffffff67b803ull, 0xffffffffff67b910ull, 0xfffffff67b138918ull,
    0xffffffffff7b621aull, 0xfffffff7b6803a21ull, 0xfffffff7b69a2092ull, 0xffff89a38a3a27b6ull,
    0xffffffffff726327ull, 0xfffffff026067807ull, 0xfffffff910732672ull, 0xffff678891681261ull,
    0xfffffff73171a67aull, 0xffff801781a7167aull, 0xffff7a69a0a70730ull, 0xfffffff9a88a7a67ull,
    0xffffffffff68b486ull, 0xfffffff640603b63ull, 0xfffffff109648b68ull, 0xffff63b139369649ull,
    0xfffffff1a28b6486ull, 0xffff640b60b03a21ull, 0xffff9a2920b648b4ull, 0xf36463b34923a39aull,
    0xfffffff264248328ull, 0xffffffffff264240ull, 0xffff834642432091ull, 0xfffffff642241491ull,
    0xffff1a6648168318ull, 0xfffffff40660a01aull, 0xf39a9303a6834364ull, 0xffffffffff4a649aull,
    0xffffffffffb67594ull, 0xfffffff67b594380ull, 0xfffffffb67045105ull, 0xffff51345343867bull,
    0xfffffffb6721a459ull, 0xffff594380a217b6ull, 0xffff204a24a45b67ull, 0xf67b25a523453843ull,
    0xfffffff945267327ull, 0xffff786260680459ull, 0xffff045051673263ull, 0xf851584812786826ull,
    0xffff73167161a459ull, 0xf459078701671a61ull, 0xfa737a6a305a4a04ull, 0xffffa84a458a7a67ull,
    0xfffffff98b9b6596ull, 0xffff590650360b63ull, 0xffffb65510b508b0ull, 0xfffffff1355363b6ull,
    0xffff65b8b9b59a21ull, 0xfa21965690b603b0ull, 0xf52025a50865b58bull, 0xffff35a3a25363b6ull,
    0xffff283265825985ull, 0xfffffff260069659ull, 0xf826283865081851ull, 0xffffffffff612651ull,
    0xf698965683a61631ull, 0xffff06505960a01aull, 0xffffffffffa65830ull, 0xfffffffffffff65aull,
    0xffffffffffb57a5bull, 0xfffffff03857ba5bull, 0xfffffff091ba57b5ull, 0xffff1381897ba57aull,
    0xfffffff15717b21bull, 0xffffb27571721380ull, 0xffff7b2209729579ull, 0xf289823295b27257ull,
    0xfffffff573532a52ull, 0xffff52a578258028ull, 0xffff2a37353a5109ull, 0xf25752a278129289ull,
    0xffffffffff573531ull, 0xfffffff571170780ull, 0xfffffff735539309ull, 0xffffffffff795789ull,
    0xfffffff8ba8a5485ull, 0xffff03bba50b5405ull, 0xffff54aba8a48910ull, 0xf41314943b54a4baull,
    0xffff8548b2582152ull, 0xfb151b2b543b0b40ull, 0xf58b8545b2950520ull, 0xffffffffff3b2549ull,
    0xffff483543253a52ull, 0xfffffff0244252a5ull, 0xf910854583a532a3ull, 0xffff2492914252a5ull,
    0xfffffff153358548ull, 0xffffffffff501540ull, 0xffff530509358548ull, 0xfffffffffffff549ull,
    0xfffffffba9b947b4ull, 0xffffba97b9794380ull, 0xffffb470414b1ba1ull, 0xf4bab474a1843413ull,
    0xffff219b294b97b4ull, 0xf3801b2b197b9479ull, 0xfffffff04224b47bull, 0xffff42343824b47bull,
    0xffff947732972a92ull, 0xf70207872a4797a9ull, 0xfa040a1a472a3a73ull, 0xffffffffff4782a1ull,
    0xfffffff317714194ull, 0xffff178180714194ull, 0xffffffffff347304ull, 0xfffffffffffff784ull,
    0xffffffffff8ba8a9ull, 0xfffffffa9bb93903ull, 0xfffffffba88a0a10ull, 0xffffffffffa3ba13ull,
    0xfffffff8b99b1b21ull, 0xffff9b2921b93903ull, 0xffffffffffb08b20ull, 0xfffffffffffffb23ull,
    0xfffffff98aa82832ull, 0xffffffffff2902a9ull, 0xffff8a1810a82832ull, 0xfffffffffffff2a1ull,
    0xffffffffff819831ull, 0xfffffffffffff190ull, 0xfffffffffffff830ull, 0xffffffffffffffffull,
};

// cube corners in the reference's naming pXYZ (offset +1 along x / y / z), in the order of its distArray; the two ends of
// cube edge e in the order vertexInterp receives them (:216-227); the bit a corner contributes to the case index (:171-178)
__constant__ int kEdgeA[12] = {2, 4, 1, 0, 5, 7, 6, 3, 2, 4, 1, 0};      // indices into {000,100,010,001,110,011,101,111}
__constant__ int kEdgeB[12] = {4, 1, 0, 2, 7, 6, 3, 5, 5, 7, 6, 3};
__constant__ int kCornerBit[8] = {8, 4, 1, 128, 2, 16, 64, 32};
__constant__ int kCornerOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};

constexpr float kInvalid = __builtin_nanf("");

// corner lattice: (X+1) x (Y+1) x (Z+1) points; corner (i,j,k) averages voxels (i-1..i, j-1..j, k-1..k)
__global__ void __launch_bounds__(256) mc_corner_kernel(const float* __restrict__ vol, int X, int Y, int Z, float truncation,
                                                        float* __restrict__ corner) {
    const long long n = (long long)(X + 1) * (Y + 1) * (Z + 1);
    const long long t = blockIdx.x * 256ll + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % (Z + 1)); const long long u = t / (Z + 1);
    const int j = (int)(u % (Y + 1)), i = (int)(u / (Y + 1));
    float out = kInvalid;
    if (i >= 1 && i < X && j >= 1 && j < Y && k >= 1 && k < Z) {
        // trilerp's order: (0,0,0) (1,0,0) (0,1,0) (0,0,1) (1,1,0) (0,1,1) (1,0,1) (1,1,1), offsets relative to voxel (i-1, j-1, k-1)
        const int ox[8] = {0, 1, 0, 0, 1, 0, 1, 1}, oy[8] = {0, 0, 1, 0, 1, 1, 0, 1}, oz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
        float dist = 0.f; bool ok = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float d = __ldg(vol + ((long long)(i - 1 + ox[c]) * Y + (j - 1 + oy[c])) * Z + (k - 1 + oz[c]));
            ok = ok && (d != -INFINITY) && (fabsf(d) < truncation);          // NaN fails the comparison, like the reference
            dist = __fadd_rn(dist, __fmul_rn(0.125f, d));
        }
        if (ok) out = dist;
    }
    corner[t] = out;
}

__device__ __forceinline__ float3 vertex_interp(float iso, float3 p1, float3 p2, float d1, float d2, int& snapped) {
    snapped = 0;
    if (fabsf(iso - d1) < 0.00001f) { snapped = 1; return p1; }
    if (fabsf(iso - d2) < 0.00001f) { snapped = 2; return p2; }
    if (fabsf(d1 - d2) < 0.00001f) { snapped = 1; return p1; }
    const float mu = __fdiv_rn(__fsub_rn(iso, d1), __fsub_rn(d2, d1));
    return make_float3(__fadd_rn(p1.x, __fmul_rn(mu, __fsub_rn(p2.x, p1.x))), __fadd_rn(p1.y, __fmul_rn(mu, __fsub_rn(p2.y, p1.y))),
                       __fadd_rn(p1.z, __fmul_rn(mu, __fsub_rn(p2.z, p1.z))));
}

// One thread per lattice point (i, j, k), k fastest.  EMIT = false: counts[t] = triangles of the cell.  EMIT = true:
// offsets[t] = first triangle of the cell; tris [T][3][3] positions in voxel units; keys [T][3] = identity of each vertex
// for welding: 4 * (corner-lattice index of the lower end of its dual edge) + axis, or 4 * corner index + 3 when the
// vertex was snapped onto a corner.
template <bool EMIT>
__global__ void __launch_bounds__(256) mc_cell_kernel(const float* __restrict__ corner, int X, int Y, int Z, float iso, float thresh,
                                                      int* __restrict__ counts, const long long* __restrict__ offsets,
                                                      float* __restrict__ tris, long long* __restrict__ keys) {
    const long long n = (long long)X * Y * Z;
    const long long t = blockIdx.x * 256ll + threadIdx.x;
    if (t >= n) return;
    const int k = (int)(t % Z); const long long u = t / Z;
    const int j = (int)(u % Y), i = (int)(u / Y);
    float d[8]; long long cid[8];
    bool valid = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        cid[c] = ((long long)(i + kCornerOff[c][0]) * (Y + 1) + (j + kCornerOff[c][1])) * (Z + 1) + (k + kCornerOff[c][2]);
        d[c] = __ldg(corner + cid[c]);
        valid = valid && (d[c] == d[c]);
    }
    int ntri = 0;
    unsigned long long row = ~0ull;
    if (valid) {
        unsigned cube = 0;
#pragma unroll
        for (int c = 0; c < 8; ++c) if (d[c] < iso) cube += kCornerBit[c];
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            ok = ok && !(fabsf(d[a]) > thresh);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (d[a] * d[b] < 0.0f) ok = ok && !(fabsf(d[a]) + fabsf(d[b]) > thresh);
                else ok = ok && !(fabsf(d[a] - d[b]) > thresh);
            }
        }
        if (ok) {
            row = kTriTable[cube];
            while (ntri < 5 && ((row >> (12 * ntri)) & 0xf) != 0xf) ++ntri;
            unsigned edges = 0;                                  // the reference's edgeTable entry = the set of edges the case uses
            for (int q = 0; q < 3 * ntri; ++q) edges |= 1u << ((row >> (4 * q)) & 0xf);
            if (edges == 255u) ntri = 0;                         // cases whose edge set is exactly 0..7 are dropped by the reference (:204)
        }
    }
    if (!EMIT) { counts[t] = ntri; return; }
    if (ntri == 0) return;
    const float3 pos = make_float3((float)i, (float)j, (float)k);
    long long o = offsets[t];
    for (int tr = 0; tr < ntri; ++tr, ++o) {
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            const int e = (int)((row >> (12 * tr + 4 * v)) & 0xf);
            const int a = kEdgeA[e], b = kEdgeB[e];
            const float3 pa = make_float3(pos.x + (kCornerOff[a][0] ? 0.5f : -0.5f), pos.y + (kCornerOff[a][1] ? 0.5f : -0.5f), pos.z + (kCornerOff[a][2] ? 0.5f : -0.5f));
            const float3 pb = make_float3(pos.x + (kCornerOff[b][0] ? 0.5f : -0.5f), pos.y + (kCornerOff[b][1] ? 0.5f : -0.5f), pos.z + (kCornerOff[b][2] ? 0.5f : -0.5f));
            int snapped;
            const float3 p = vertex_interp(iso, pa, pb, d[a], d[b], snapped);
            tris[9 * o + 3 * v] = p.x; tris[9 * o + 3 * v + 1] = p.y; tris[9 * o + 3 * v + 2] = p.z;
            long long key;
            if (snapped) key = 4 * (snapped == 1 ? cid[a] : cid[b]) + 3;
            else {
                const int axis = (kCornerOff[a][0] != kCornerOff[b][0]) ? 0 : (kCornerOff[a][1] != kCornerOff[b][1]) ? 1 : 2;
                key = 4 * (cid[a] < cid[b] ? cid[a] : cid[b]) + axis;
            }
            keys[3 * o + v] = key;
        }
    }
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" int64_t rf_mc_corner_floats(int X, int Y, int Z) { return (X < 1 || Y < 1 || Z < 1) ? 0 : (int64_t)(X + 1) * (Y + 1) * (Z + 1); }

extern "C" int rf_mc_count(const float* volume, int X, int Y, int Z, float isovalue, float truncation, float* corner_ws, int* counts, void* stream) {
    RF_REQUIRE(volume && corner_ws && counts, RF_E_NULL, "rf_mc_count: NULL pointer");
    RF_REQUIRE(X >= 1 && Y >= 1 && Z >= 1 && (long long)(X + 1) * (Y + 1) * (Z + 1) < (1ll << 40), RF_E_RANGE, "rf_mc_count: bad volume dimensions");
    cudaStream_t s = (cudaStream_t)stream;
    const long long nc = (long long)(X + 1) * (Y + 1) * (Z + 1), n = (long long)X * Y * Z;
    mc_corner_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(volume, X, Y, Z, truncation, corner_ws);
    RF_CHECK_LAUNCH("mc_corner_kernel");
    mc_cell_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(corner_ws, X, Y, Z, isovalue, 10.0f, counts, nullptr, nullptr, nullptr);
    RF_CHECK_LAUNCH("mc_cell_kernel<count>");
    return 0;
}

extern "C" int rf_mc_emit(const float* corner_ws, int X, int Y, int Z, float isovalue, const long long* offsets, float* triangles,
                          long long* keys, void* stream) {
    RF_REQUIRE(corner_ws && offsets && triangles && keys, RF_E_NULL, "rf_mc_emit: NULL pointer");
    RF_REQUIRE(X >= 1 && Y >= 1 && Z >= 1, RF_E_RANGE, "rf_mc_emit: bad volume dimensions");
    const long long n = (long long)X * Y * Z;
    mc_cell_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(corner_ws, X, Y, Z, isovalue, 10.0f, nullptr, offsets, triangles, keys);
    RF_CHECK_LAUNCH("mc_cell_kernel<emit>");
    return 0;
}
