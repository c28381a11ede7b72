// Internal C++ launch API of libvog_b200 (one function per kernel family).  The exported C ABI in
// vog_abi.cu (declared in include/vog_b200.h) is a thin argument-checking shim over these.
#pragma once
#include <cuda_runtime.h>

#define VOG_MAX_HEADS 8

namespace vog {

// ---- fp32_path.cu : exact-fp32 CUDA-core kernels --------------------------------------------
int sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias,
             const float* R, int ldr, float* C, int ldc, int M, int N, int K, int relu,
             cudaStream_t st);
int attn_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo,
             int Bt, int N, int H, const int* off, const int* dh, float inv_scale,
             int bias_mode, const float* a, int nbox, const float* bpe, const float* dense,
             cudaStream_t st);
int add_layernorm(const float* x, int ldx, const float* r, int ldr, const float* w, const float* b,
                  float* out, int ldo, void* out_lp, int ldlp, int lp_kind, int M, int d, float eps,
                  cudaStream_t st);
int pe_project(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw,
               float vh, float fdiv, float scale, cudaStream_t st);
int select_fwd(const float* scores, const float* props, int pdim, float* boxes, float* out_scores,
               long long* indexs, int B, int nsrl, int ncmp, int nfrm, int nppf, int spat,
               cudaStream_t st);

}  // namespace vog
