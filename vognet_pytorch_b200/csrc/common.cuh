// Shared host/device helpers for libvog_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace vog {

// ---- error plumbing (C ABI: int return + vog_last_error()) ---------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);     // cudaGetLastError() -> 0 / -1 (+message)

#define VOG_REQUIRE(cond, ...)                          \
    do {                                                \
        if (!(cond)) {                                  \
            ::vog::set_error(__VA_ARGS__);              \
            return -1;                                  \
        }                                               \
    } while (0)

#define VOG_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            ::vog::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),     \
                             __FILE__, __LINE__);                                        \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return cdiv(a, b) * b; }

// ---- small device helpers ------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace vog
