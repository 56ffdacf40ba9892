// umma.cuh — minimal hand-written tcgen05 (5th-gen tensor core) toolkit for sm_100a: TMEM allocation, shared-memory
// matrix descriptors for the un-swizzled canonical layouts, single-thread MMA issue, commit -> mbarrier, TMEM loads.
//
// Operand layout used throughout ("chunked", no swizzle): a [rows x K] bf16 operand is stored as 16-byte chunks of 8
// consecutive K elements; chunk c of row r lives at   c * (rows*16) + (r/8)*128 + (r%8)*16   bytes.  This is the
// canonical SWIZZLE_NONE layout for BOTH majors:
//   * K-major use (row = M or N index, contraction over the chunked dimension):  LBO = rows*16, SBO = 128;
//   * MN-major use (chunked dimension = M or N index, contraction over rows):    LBO = 128,     SBO = rows*16;
// so the same buffer serves Y = X * W^T (contraction over features) and dW^T = X^T * dY (contraction over samples).
// A thread that owns row r writes one 16-byte vector per chunk; a warp's 32 rows are 512 contiguous bytes.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace rf {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM ------------------------------------------------------------------------------------------------
// one full warp executes alloc/dealloc; ncols power of two >= 32
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a fully converged warp (elect.sync): ptxas recognises the pattern and issues the guarded tcgen05 instructions once,
// from uniform registers, instead of wrapping each of them in a loop over the active lanes.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: returns false on timeout (a wrong descriptor must not hang the GPU)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int i = 0; i < (1 << 22); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): global -> shared, completion counted in bytes on an mbarrier ---------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 16-byte aligned source / destination, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// 32-bit instruction descriptor for kind::f16 with bf16 inputs, fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t idesc_bf16(int n, bool a_mn_major, bool b_mn_major) {
    return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// advance the start-address field of a descriptor by `bytes` (low word only: the field never carries out, smem < 256 KB)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) {
    return (d & 0xFFFFFFFF00000000ull) | (uint64_t)((uint32_t)d + (bytes >> 4));
}

// D[tmem] (+)= A * B^T ; one thread issues
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B^T : A read from tensor memory (lane = row, 32-bit column c = K elements 2c, 2c+1 as a bf16 pair,
// K-major only), B from shared memory.  No shared-memory bandwidth is spent on A.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM (thread i of warp w writes lane 32*(w%4)+i; 4 or 8 consecutive 32-bit columns)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// arrive on the mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- chunked operand layout ------------------------------------------------------------------------------
// byte offset of chunk c (8 bf16) of row r in a [rows x *] operand
__device__ __forceinline__ uint32_t chunk_off(int rows, int r, int c) { return (uint32_t)(c * rows * 16 + (r >> 3) * 128 + (r & 7) * 16); }
// K-major descriptor for MMA k-step starting at chunk c0 (covers chunks c0, c0+1), rows starting at row r0 (multiple of 8)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int rows, int c0, int r0 = 0) {
    return smem_desc(base + (uint32_t)(c0 * rows * 16 + (r0 >> 3) * 128), (uint32_t)(rows * 16), 128u);
}
// MN-major descriptor: M/N index = chunked dimension starting at chunk c0; contraction over rows r0 .. r0+15
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int rows, int c0, int r0) {
    return smem_desc(base + (uint32_t)(c0 * rows * 16 + (r0 >> 3) * 128), 128u, (uint32_t)(rows * 16));
}

// split fp32 into bf16 hi + bf16 lo (x ~= hi + lo, relative error ~2^-17)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// pack 8 floats into one 16-byte chunk of hi parts and one of lo parts.  cvt.rn.bf16x2.f32 d, a, b packs a into the
// upper and b into the lower half: two elements per conversion, 6 instructions per pair for the whole split.
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
        const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l[i]) : "f"(x[2 * i + 1] - h1), "f"(x[2 * i] - h0));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- TMEM -> registers (thread i of warp w reads lane 32*(w%4)+i; N consecutive fp32 columns) ---------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
                 "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma
}  // namespace rf
