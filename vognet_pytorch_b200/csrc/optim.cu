// Fused Adam step on a flat fp32 buffer ("Adam(0.9, 0.99) on the flat buffer right after
// the all-reduce").  The reference trains with torch.optim.Adam(betas=(0.9, 0.99)), lr 1e-4, eps 1e-8, no weight decay
// (code/main_dist.py:55, utils/trn_utils.py:799-803, configs/anet_srl_cfg.yml:108): ~10 elementwise library kernels per
// parameter tensor and step (x 110 tensors).  Here ONE launch updates all 45 M parameters: 16 B read + 12 B written per
// element, HBM-bound.  Arithmetic follows torch's single-tensor Adam operation by operation in fp32:
//     m  = fma(1 - b1, g - m, m)                              (lerp_)
//     v  = fma((1 - b2) * g, g, v * b2)                       (mul_ + addcmul_)
//     p  = fma(-(lr / bc1), m / (sqrt(v) / sqrt(bc2) + eps), p)   (addcdiv_)
// with the bias corrections bc1 = 1 - b1^t, bc2 = 1 - b2^t evaluated in double on the host like torch does in Python.
// `grad_scale` folds the 1/world of a summed gradient all-reduce into the same pass.
#include <math.h>
#include "common.cuh"
#include "kernels.h"

namespace vog {

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float w1, float b2, float w2, float step_size,
                                      float bc2_sqrt, float eps, float gs)
{
    g = __fmul_rn(g, gs);
    // the a + b * c forms are single fused multiply-adds, as in torch's own elementwise kernels (lerp, addcmul, addcdiv
    // are compiled with floating-point contraction on)
    m = fmaf(w1, __fsub_rn(g, m), m);
    v = fmaf(__fmul_rn(w2, g), g, __fmul_rn(v, b2));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = fmaf(-step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long n, float w1, float b2, float w2, float step_size, float bc2_sqrt, float eps, float gs)
{
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
        adam1(P.x, G.x, M.x, V.x, w1, b2, w2, step_size, bc2_sqrt, eps, gs);
        adam1(P.y, G.y, M.y, V.y, w1, b2, w2, step_size, bc2_sqrt, eps, gs);
        adam1(P.z, G.z, M.z, V.z, w1, b2, w2, step_size, bc2_sqrt, eps, gs);
        adam1(P.w, G.w, M.w, V.w, w1, b2, w2, step_size, bc2_sqrt, eps, gs);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {                    // tail (n not a multiple of 4)
        const long long i = (n4 << 2) + threadIdx.x;
        adam1(p[i], g[i], m[i], v[i], w1, b2, w2, step_size, bc2_sqrt, eps, gs);
    }
}

int adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
              long long step, double grad_scale, cudaStream_t st)
{
    VOG_REQUIRE(n >= 0 && step >= 1, "adam_step: need n >= 0 and step >= 1 (got %lld, %lld)", n, step);
    if (n == 0) return 0;
    VOG_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
    // bias corrections 1 - beta^t exactly as torch.optim.Adam computes them (beta ** step in double precision)
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    const long long n4 = n >> 2;
    long long blocks = (n4 + 255) / 256;
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                                      (float)(lr / bc1), (float)sqrt(bc2), (float)eps, (float)grad_scale);
    return check_launch("adam_step");
}

}  // namespace vog
