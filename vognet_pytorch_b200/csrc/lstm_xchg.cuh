// Cross-CTA exchange primitives of the persistent LSTM kernels (lstm_rec.cu forward, lstm_bwd.cu backward): relaxed
// gpu-scope stores / loads that bypass the non-coherent L1, and word groups that validate themselves through a tag.
#pragma once
#include <cuda_runtime.h>

namespace vog {

__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 ld_relaxed_u64x2(const unsigned long long* p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <int VW> struct XVec { unsigned w[VW]; };
template <int VW>
__device__ __forceinline__ XVec<VW> ld_relaxed_words(const unsigned* p) {        // VW = 1, 2, 4 words, naturally aligned
    XVec<VW> v;
    if constexpr (VW == 4)
        asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]) : "l"(p) : "memory");
    else if constexpr (VW == 2)
        asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.w[0]), "=r"(v.w[1]) : "l"(p) : "memory");
    else
        asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v.w[0]) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_u32x4(unsigned* p, unsigned a, unsigned b, unsigned c, unsigned d) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

}  // namespace vog
