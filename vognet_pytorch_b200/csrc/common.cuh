// Shared host/device helpers for libvog_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>

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

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// The forward at the gt5 sizes is a chain of ~25 short dependent kernels: each one's launch latency and on-chip set-up
// (barrier init, TMEM allocation, tensor-map prefetch, weight loads) would otherwise sit on the critical path.  Kernels
// launched through launch_pdl() may START as soon as every CTA of the preceding kernel of the stream has executed
// pdl_trigger(); they must call pdl_wait() before their first access to memory another kernel of the stream produces
// (it returns once the preceding grid has completed and its writes are visible).  Only data written by plain launches
// (parameters, packed weights) may be read before pdl_wait().  Both are no-ops under a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();             // thread-local switch (vog_debug_pdl / VOG_PDL=1), default OFF: see DESIGN.md section 6.0
void pdl_set(bool on);

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
