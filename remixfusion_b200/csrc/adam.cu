// adam.cu — N1 of SURVEY §8(f): the optimiser step that follows every backward of the mapping loop.
//
// Replaces torch.optim.Adam as the reference configures it (mp_slam/slam.py:271-286: betas (0.9, 0.99); decoder group
// weight_decay 1e-6; hash-table group eps 1e-15) and the zero_grad that follows it (mp_slam/mapper.py:417-423) with one
// pass over (param, grad, exp_avg, exp_avg_sq): 16 B read + 12 B written per parameter (+4 B when the gradient is
// cleared in the same pass) instead of the ~10 elementwise passes of the unfused optimiser.  Semantics are
// torch's _single_tensor_adam, dense: an entry whose gradient is zero still moves by momentum and weight decay.
//   g   = grad + weight_decay * p
//   m   = m + (g - m) * (1 - beta1)                      (lerp)
//   v   = v * beta2 + (1 - beta2) * g * g                (mul, addcmul)
//   p   = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
#include <math.h>
#include <algorithm>
#include "rf_common.cuh"

namespace rf {

struct AdamK { float w1, beta2, w2, wd, eps, step_size, inv_bc2_sqrt_div; int zero_grad; };
struct AdamDyn { double lr, beta1, beta2; const float* step; };      // step count read on the device (CUDA-graph capture)

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamK& k) {
    float gg = (k.wd != 0.f) ? __fadd_rn(g, __fmul_rn(k.wd, p)) : g;
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(gg, m), k.w1));
    v = __fadd_rn(__fmul_rn(v, k.beta2), __fmul_rn(__fmul_rn(k.w2, gg), gg));
    float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), k.inv_bc2_sqrt_div), k.eps);
    p = __fadd_rn(p, __fmul_rn(-k.step_size, __fdiv_rn(m, denom)));
    if (k.zero_grad) g = 0.f;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   long long n, AdamK k, AdamDyn dyn) {
    if (dyn.step) {                                        // same float64 scalars the host path derives, from the device step count
        const double t = (double)__ldg(dyn.step);
        const double bc1 = 1.0 - pow(dyn.beta1, t), bc2 = 1.0 - pow(dyn.beta2, t);
        k.step_size = (float)(dyn.lr / bc1); k.inv_bc2_sqrt_div = (float)sqrt(bc2);
    }
    const long long n4 = n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        adam_one(P.x, G.x, M.x, V.x, k); adam_one(P.y, G.y, M.y, V.y, k); adam_one(P.z, G.z, M.z, V.z, k); adam_one(P.w, G.w, M.w, V.w, k);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
        if (k.zero_grad) reinterpret_cast<float4*>(g)[i] = G;
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        float P = p[i], G = g[i], M = m[i], V = v[i];
        adam_one(P, G, M, V, k);
        p[i] = P; m[i] = M; v[i] = V;
        if (k.zero_grad) g[i] = G;
    }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                            double beta2, double eps, double weight_decay, int64_t step, const float* step_dev, int zero_grad, void* stream) {
    RF_REQUIRE(n >= 0 && (step >= 1 || step_dev), RF_E_RANGE, "rf_adam_step: n >= 0 and step >= 1 (or a device step count) required");
    if (step < 1) step = 1;
    if (n == 0) return 0;
    RF_REQUIRE(param && grad && exp_avg && exp_avg_sq, RF_E_NULL, "rf_adam_step: NULL pointer");
    RF_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, RF_E_ALIGN, "rf_adam_step: 16-byte alignment");
    // scalars exactly as torch derives them in Python floats (float64), narrowed where torch narrows them
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    AdamK k;
    k.w1 = (float)(1.0 - beta1); k.beta2 = (float)beta2; k.w2 = (float)(1.0 - beta2); k.wd = (float)weight_decay;
    k.eps = (float)eps; k.step_size = (float)(lr / bc1); k.inv_bc2_sqrt_div = (float)sqrt(bc2); k.zero_grad = zero_grad ? 1 : 0;
    long long n4 = (n + 3) / 4;
    int blocks = (int)std::min<long long>((n4 + 255) / 256, (long long)num_sms() * 8);
    AdamDyn dyn{lr, beta1, beta2, step_dev};
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (long long)n, k, dyn);
    RF_CHECK_LAUNCH("adam_kernel");
    return 0;
}
