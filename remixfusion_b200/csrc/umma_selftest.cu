// umma_selftest.cu — hardware check of the tcgen05 building blocks in umma.cuh (descriptors, TMEM, commit/wait):
//   mode 0:  D[128][N] = A[128][K] * B[N][K]^T      (both operands K-major)
//   mode 1:  D[f][j]   = sum_m X[m][f] * Y[m][j]     (both operands MN-major: contraction over the 128 rows)
//   mode 2:  D[128][N] = A[128][K] * W[K][N]         (A K-major; W stored with K rows and used MN-major: the form the
//            decoder backward uses to multiply by a weight matrix kept in its forward layout), K in {16,32,64}
//   mode 3:  as mode 0 with A read from TENSOR MEMORY (tcgen05.mma TS form): each thread stores its row as bf16 pairs
// bf16x3 split (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM.  One CTA of 128 threads.
#include "rf_common.cuh"
#include "umma.cuh"

namespace rf {

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D,
                                                            int K, int N, int mode, int* status) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    // mode 0: A [128 x K] (K/8 chunks, 128 rows), B [N x K] (K/8 chunks, N rows)
    // mode 1: X [128 x K] used as A MN-major (needs 16 feature chunks -> zero padded), Y [128 x N] (N/8 chunks, 128 rows)
    const int a_chunks = (mode == 1) ? 16 : K / 8;
    const int b_rows = (mode == 0 || mode == 3) ? N : (mode == 1) ? 128 : K;
    const int b_chunks = (mode == 0 || mode == 3) ? K / 8 : N / 8;
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + a_chunks * 128 * 16;
    unsigned char* b_hi = a_lo + a_chunks * 128 * 16;
    unsigned char* b_lo = b_hi + b_chunks * b_rows * 16;
    if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    // stage operands
    for (int c = 0; c < a_chunks; ++c) {
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = (c * 8 + i < K) ? A[tid * K + c * 8 + i] : 0.f;
        uint4 h, l; umma::split8(v, h, l);
        *reinterpret_cast<uint4*>(a_hi + umma::chunk_off(128, tid, c)) = h;
        *reinterpret_cast<uint4*>(a_lo + umma::chunk_off(128, tid, c)) = l;
    }
    if (tid < b_rows) {
        const int bk = (mode == 0 || mode == 3) ? K : N;
        for (int c = 0; c < b_chunks; ++c) {
            float v[8];
            for (int i = 0; i < 8; ++i) v[i] = B[tid * bk + c * 8 + i];
            uint4 h, l; umma::split8(v, h, l);
            *reinterpret_cast<uint4*>(b_hi + umma::chunk_off(b_rows, tid, c)) = h;
            *reinterpret_cast<uint4*>(b_lo + umma::chunk_off(b_rows, tid, c)) = l;
        }
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tb = tmem_base;
    if (mode == 3) {
        // A hi at TMEM columns [64, 64 + K/2), A lo at [128, 128 + K/2): one 32-bit column per bf16 pair (k, k+1)
        const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < K / 8; ++c) {
            float v[8];
            for (int i = 0; i < 8; ++i) v[i] = A[tid * K + c * 8 + i];
            uint4 h, l; umma::split8(v, h, l);
            umma::tmem_st4(tl + 64 + 4 * c, h);
            umma::tmem_st4(tl + 128 + 4 * c, l);
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }
    if (tid == 0) {
        const uint32_t ah = umma::smem_u32(a_hi), al = umma::smem_u32(a_lo), bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
        uint32_t acc = 0;
        if (mode == 3) {
            const uint32_t id = umma::idesc_bf16(N, false, false);
            for (int s = 0; s < K / 16; ++s) {
                uint64_t dbh = umma::desc_kmajor(bh, N, 2 * s), dbl = umma::desc_kmajor(bl, N, 2 * s);
                umma::mma_bf16_ts(tb, tb + 64 + 8 * s, dbh, id, acc); acc = 1;
                umma::mma_bf16_ts(tb, tb + 64 + 8 * s, dbl, id, 1);
                umma::mma_bf16_ts(tb, tb + 128 + 8 * s, dbh, id, 1);
            }
        } else if (mode == 0) {
            const uint32_t id = umma::idesc_bf16(N, false, false);
            for (int s = 0; s < K / 16; ++s) {
                uint64_t dah = umma::desc_kmajor(ah, 128, 2 * s), dal = umma::desc_kmajor(al, 128, 2 * s);
                uint64_t dbh = umma::desc_kmajor(bh, N, 2 * s), dbl = umma::desc_kmajor(bl, N, 2 * s);
                umma::mma_bf16(tb, dah, dbh, id, acc); acc = 1;
                umma::mma_bf16(tb, dah, dbl, id, 1);
                umma::mma_bf16(tb, dal, dbh, id, 1);
            }
        } else if (mode == 2) {
            const uint32_t id = umma::idesc_bf16(N, false, true);
            for (int s = 0; s < K / 16; ++s) {
                uint64_t dah = umma::desc_kmajor(ah, 128, 2 * s), dal = umma::desc_kmajor(al, 128, 2 * s);
                uint64_t dbh = umma::desc_mnmajor(bh, K, 0, 16 * s), dbl = umma::desc_mnmajor(bl, K, 0, 16 * s);
                umma::mma_bf16(tb, dah, dbh, id, acc); acc = 1;
                umma::mma_bf16(tb, dah, dbl, id, 1);
                umma::mma_bf16(tb, dal, dbh, id, 1);
            }
        } else {
            const uint32_t id = umma::idesc_bf16(N, true, true);
            for (int s = 0; s < 8; ++s) {
                uint64_t dah = umma::desc_mnmajor(ah, 128, 0, 16 * s), dal = umma::desc_mnmajor(al, 128, 0, 16 * s);
                uint64_t dbh = umma::desc_mnmajor(bh, 128, 0, 16 * s), dbl = umma::desc_mnmajor(bl, 128, 0, 16 * s);
                umma::mma_bf16(tb, dah, dbh, id, acc); acc = 1;
                umma::mma_bf16(tb, dah, dbl, id, 1);
                umma::mma_bf16(tb, dal, dbh, id, 1);
            }
        }
        umma::commit(&bar);
    }
    bool ok = umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (!ok) { if (tid == 0) *status = 1; }
    else {
        const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 16) {
            float v[16];
            umma::tmem_ld16(taddr + n0, v);
            for (int i = 0; i < 16; ++i) D[tid * N + n0 + i] = v[i];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 256);
}

}  // namespace rf

// A: [128,K] fp32, B: [N,K] (modes 0, 3), [128,N] (mode 1) or [K,N] (mode 2), D: [128,N]; returns 0 ok, 1 = MMA never completed (timeout)
extern "C" int rf_umma_selftest(const float* A, const float* B, float* D, int K, int N, int mode, void* stream) {
    RF_REQUIRE(A && B && D, RF_E_NULL, "rf_umma_selftest: NULL");
    RF_REQUIRE(N >= 16 && N <= 64 && N % 16 == 0 && K >= 16 && K <= 128 && K % 16 == 0 && mode >= 0 && mode <= 3, RF_E_RANGE, "rf_umma_selftest: bad shape");
    int* status; cudaMalloc(&status, sizeof(int)); cudaMemset(status, 0, sizeof(int));
    size_t sm = 2 * 16 * 128 * 16 + 2 * 16 * 128 * 16 + 1024;
    cudaFuncSetAttribute(rf::umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    rf::umma_selftest_kernel<<<1, 128, sm, (cudaStream_t)stream>>>(A, B, D, K, N, mode, status);
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    int h = 0; cudaMemcpy(&h, status, sizeof(int), cudaMemcpyDeviceToHost); cudaFree(status);
    if (e != cudaSuccess) return rf::set_error((int)e, "rf_umma_selftest: %s", cudaGetErrorString(e));
    if (h) return rf::set_error(1000 + h, "rf_umma_selftest: tensor-core pipeline timed out (status %d)", h);
    return 0;
}
