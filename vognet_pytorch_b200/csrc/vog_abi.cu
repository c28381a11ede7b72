// extern "C" surface of libvog_b200 (declared in include/vog_b200.h).
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "../../include/vog_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace vog {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
static thread_local bool g_pdl = false;         // the model switches it on while it captures the forward of a small configuration
bool pdl_enabled() { return g_pdl; }
void pdl_set(bool on) { g_pdl = on; }

int check_launch(const char* what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

}  // namespace vog

using namespace vog;

extern "C" {

const char* vog_last_error(void) { return g_err; }
int vog_abi_version(void) { return 4; }
long long vog_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
void vog_debug_pdl(int on) { vog::pdl_set(on != 0); }

int vog_device_is_sm100(void)
{
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10;
}

int vog_sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias,
                 const float* R, int ldr, float* C, int ldc, int M, int N, int K, int relu,
                 void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0 && K >= 0, "vog_sgemm_nt: negative dimension");
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(A && W && C, "vog_sgemm_nt: null operand");
    VOG_REQUIRE(lda >= K && ldw >= K && ldc >= N && (!R || ldr >= N), "vog_sgemm_nt: bad leading dimension");
    return sgemm_nt(A, lda, W, ldw, bias, R, ldr, C, ldc, M, N, K, relu, (cudaStream_t)stream);
}

int vog_attn_fwd_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo,
                     int Bt, int N, int H, const int* off, const int* dh, float inv_scale,
                     int bias_mode, const float* a, int nbox, const float* bpe,
                     const float* dense, float* lse, float drop_p, uint64_t seed, void* stream)
{
    VOG_REQUIRE(q && k && v && out && off && dh, "vog_attn_fwd_f32: null operand");
    VOG_REQUIRE(Bt >= 0 && N >= 0, "vog_attn_fwd_f32: negative dimension");
    VOG_REQUIRE(bias_mode >= 0 && bias_mode <= 2, "vog_attn_fwd_f32: bad bias_mode %d", bias_mode);
    VOG_REQUIRE(bias_mode != VOG_BIAS_RANK1 || nbox > 0, "vog_attn_fwd_f32: nbox must be > 0");
    return attn_f32(q, k, v, ld, out, ldo, Bt, N, H, off, dh, inv_scale, bias_mode, a, nbox, bpe,
                    dense, (cudaStream_t)stream, lse, drop_p, (unsigned long long)seed);
}

int vog_add_layernorm(const float* x, int ldx, const float* r, int ldr, const float* w,
                      const float* b, float* out, int ldo, void* out_lp, int ldlp, int lp_kind,
                      int M, int d, float eps, void* stream)
{
    VOG_REQUIRE(x && w && b && (out || out_lp), "vog_add_layernorm: null operand");
    VOG_REQUIRE(!out_lp || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_add_layernorm: bad lp_kind");
    return add_layernorm(x, ldx, r, ldr, w, b, out, ldo, out_lp, ldlp, lp_kind, M, d, eps,
                         (cudaStream_t)stream);
}

int vog_pe_project(const float* props, int ldp, const float* W, float* a, int rows, int H,
                   float vw, float vh, float fdiv, float scale, void* stream)
{
    VOG_REQUIRE(props && W && a, "vog_pe_project: null operand");
    VOG_REQUIRE(ldp >= 5, "vog_pe_project: proposals need >= 5 columns");
    return pe_project(props, ldp, W, a, rows, H, vw, vh, fdiv, scale, (cudaStream_t)stream);
}

int vog_pe_project_expand(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw, float vh,
                          float fdiv, float scale, int Bt, int N, int nbox, float inv_scale, void* key_factors,
                          int64_t key_factors_bytes, void* stream)
{
    VOG_REQUIRE(props && W && a && key_factors, "vog_pe_project_expand: null operand");
    VOG_REQUIRE(ldp >= 5, "vog_pe_project_expand: proposals need >= 5 columns");
    VOG_REQUIRE(Bt > 0 && N > 0 && H > 0, "vog_pe_project_expand: bad dimension");
    VOG_REQUIRE(key_factors_bytes >= vog_tc_attn_workspace_bytes(Bt, N, H) &&
                (reinterpret_cast<uintptr_t>(key_factors) & 15) == 0,
                "vog_pe_project_expand: key_factors needs vog_tc_attn_workspace_bytes(Bt, N, H) bytes, 16-byte aligned");
    return pe_project_expand(props, ldp, W, a, rows, H, vw, vh, fdiv, scale, (float*)key_factors, Bt, N, nbox,
                             tc_attn_key_ld(N), tc_attn_key_scale(inv_scale), (cudaStream_t)stream);
}

int vog_select_fwd(const float* scores, const float* props, int pdim, float* boxes,
                   float* out_scores, int64_t* indexs, int B, int nsrl, int ncmp, int nfrm,
                   int nppf, int spat, void* stream)
{
    VOG_REQUIRE(scores && props && boxes && out_scores && indexs, "vog_select_fwd: null operand");
    VOG_REQUIRE(nppf > 0 && ncmp > 0 && nfrm > 0, "vog_select_fwd: empty proposal group");
    return select_fwd(scores, props, pdim, boxes, out_scores, (long long*)indexs, B, nsrl, ncmp,
                      nfrm, nppf, spat, (cudaStream_t)stream);
}

int vog_select_sep_fwd(const float* scores, const float* props, int pdim, const float* fin_scores, float* boxes,
                       float* out_scores, int64_t* indexs, int B, int nsrl, int ncmp, int nfrm, int nppf, void* stream)
{
    VOG_REQUIRE(scores && props && fin_scores && boxes && out_scores && indexs, "vog_select_sep_fwd: null operand");
    VOG_REQUIRE(nppf > 0 && ncmp > 0 && nfrm > 0, "vog_select_sep_fwd: empty proposal group");
    return select_fwd(scores, props, pdim, boxes, out_scores, (long long*)indexs, B, nsrl, ncmp,
                      nfrm, nppf, 0, (cudaStream_t)stream, fin_scores);
}

int vog_sep_fin_scores(const float* logits, const float* vidf, const int64_t* srl_msk, const int64_t* verb_ind,
                       const int64_t* cmp_msk, float* fin_loss, float* fin_eval, int Bq, int nsrl, int P1, void* stream)
{
    VOG_REQUIRE(logits && vidf && srl_msk && verb_ind && cmp_msk && fin_loss && fin_eval, "vog_sep_fin_scores: null operand");
    return sep_fin_scores(logits, vidf, (const long long*)srl_msk, (const long long*)verb_ind, (const long long*)cmp_msk,
                          fin_loss, fin_eval, Bq, nsrl, P1, (cudaStream_t)stream);
}

int vog_verb_loss_fwd(const float* vidf, const int64_t* verb_cmp, const int64_t* verb_cross_cmp_msk, int n, int m,
                      float loss_lambda, float* loss, void* stream)
{
    VOG_REQUIRE(vidf && verb_cmp && verb_cross_cmp_msk && loss, "vog_verb_loss_fwd: null operand");
    return verb_loss_fwd(vidf, (const long long*)verb_cmp, (const long long*)verb_cross_cmp_msk, n, m, loss_lambda, loss,
                         (cudaStream_t)stream);
}

int vog_concat_videos(const float* feat, int D, const float* seg, int Ds, const float* props, int pdim, float* feat_out,
                      float* seg_out, float* props_out, int B, int ncmp, int nfrm, int nppf, int spat, float shift,
                      void* stream)
{
    return concat_videos(feat, D, seg, Ds, props, pdim, feat_out, seg_out, props_out, B, ncmp, nfrm, nppf, spat, shift,
                         (cudaStream_t)stream);
}

int vog_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, double grad_scale, void* stream)
{
    VOG_REQUIRE((param && grad && exp_avg && exp_avg_sq) || n == 0, "vog_adam_step: null operand");
    return adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

static int require_sm100(const char* who)
{
    VOG_REQUIRE(vog_device_is_sm100(), "%s: needs an sm_100 (B200) device - tcgen05/TMEM kernels have no other path", who);
    return 0;
}

int vog_cast_lp(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols, int kind,
                void* stream)
{
    VOG_REQUIRE(rows >= 0 && cols >= 0, "vog_cast_lp: negative dimension");
    if (rows == 0 || cols == 0) return 0;
    VOG_REQUIRE(src && dst, "vog_cast_lp: null operand");
    return cast_lp(src, lds, dst, ldd, rows, cols, kind, (cudaStream_t)stream);
}

int64_t vog_tc_gemm_workspace_bytes(int M, int N, int K, int tf32, int BN)
{
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return tc_gemm_workspace_bytes(M, N, K, tf32, BN);
}

int vog_tc_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int tf32,
                int BN, const float* bias, int relu, const float* residual, int64_t ldr,
                float* out_f32, int64_t ldc, void* out_lp, int64_t ldlp, int lp_kind, int rep,
                void* workspace, int64_t workspace_bytes, void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0 && K > 0, "vog_tc_gemm: bad dimension");
    if (M == 0 || N == 0) return 0;
    if (require_sm100("vog_tc_gemm")) return -1;
    VOG_REQUIRE(A && W, "vog_tc_gemm: null operand");
    TcEpilogue e;
    e.bias = bias; e.relu = relu; e.residual = residual; e.ldr = ldr;
    e.out_f32 = out_f32; e.ldc = ldc; e.out_lp = out_lp; e.ldlp = ldlp; e.lp_kind = lp_kind; e.rep = rep;
    VOG_REQUIRE(!out_lp || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_tc_gemm: bad lp_kind");
    return tc_gemm(A, lda, W, ldw, M, N, K, tf32, BN, e, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vog_tc_gemm_qkv(const void* A, int64_t lda, const void* Wqkv, int64_t ldw, int M, int K, int tf32,
                    int n_heads, int dhp, int seq_n, void* q, void* k, void* v, void* stream)
{
    VOG_REQUIRE(M >= 0 && K > 0 && n_heads >= 1 && n_heads <= VOG_MAX_HEADS, "vog_tc_gemm_qkv: bad dimension");
    if (M == 0) return 0;
    if (require_sm100("vog_tc_gemm_qkv")) return -1;
    VOG_REQUIRE(A && Wqkv && q && k && v, "vog_tc_gemm_qkv: null operand");
    VOG_REQUIRE(dhp % 64 == 0 && dhp <= 256, "vog_tc_gemm_qkv: dhp=%d must be 64/128/192/256", dhp);
    VOG_REQUIRE(seq_n > 0 && M % seq_n == 0, "vog_tc_gemm_qkv: bad sequence geometry");
    TcEpilogue e;
    e.mode = 1; e.q = (__nv_bfloat16*)q; e.k = (__nv_bfloat16*)k; e.v = (__nv_bfloat16*)v;
    e.seq_n = seq_n; e.n_heads = n_heads; e.dhp = dhp;
    return tc_gemm(A, lda, Wqkv, ldw, M, 3 * n_heads * dhp, K, tf32, dhp, e, nullptr, 0, (cudaStream_t)stream);
}

int vog_tc_gemm_qkv_factored(const void* A, int64_t lda, const void* Wvis, int64_t ldw, int M, int K, int tf32,
                             int n_heads, int dhp, const float* lq, int64_t ldq, int nfrm, int nsrl, int nppf2,
                             void* q, void* k, void* v, void* stream)
{
    VOG_REQUIRE(M >= 0 && K > 0 && n_heads >= 1 && n_heads <= VOG_MAX_HEADS, "vog_tc_gemm_qkv_factored: bad dimension");
    if (M == 0) return 0;
    if (require_sm100("vog_tc_gemm_qkv_factored")) return -1;
    VOG_REQUIRE(A && Wvis && lq && q && k && v, "vog_tc_gemm_qkv_factored: null operand");
    VOG_REQUIRE(dhp % 64 == 0 && dhp <= 256, "vog_tc_gemm_qkv_factored: dhp=%d must be 64/128/192/256", dhp);
    VOG_REQUIRE(nfrm > 0 && nsrl > 0 && nppf2 > 0, "vog_tc_gemm_qkv_factored: bad sequence geometry");
    TcEpilogue e;
    e.mode = 2; e.q = (__nv_bfloat16*)q; e.k = (__nv_bfloat16*)k; e.v = (__nv_bfloat16*)v;
    e.seq_n = nsrl * nppf2; e.n_heads = n_heads; e.dhp = dhp;
    e.lq = lq; e.ldq = ldq; e.nsrl = nsrl; e.nppf2 = nppf2; e.nfrm = nfrm;
    return tc_gemm(A, lda, Wvis, ldw, M, 3 * n_heads * dhp, K, tf32, dhp, e, nullptr, 0, (cudaStream_t)stream);
}

int vog_tc_gemm_lin2(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int tf32,
                     const float* bias, const float* w2, const float* b2, const int64_t* srl_msk,
                     const int64_t* cmp_msk, float* logits, float* scores, int B, int nfrm, int nsrl, int nppf2,
                     int ncmp, int nppf, int nfrm0, int spat, void* stream)
{
    VOG_REQUIRE(M >= 0 && N > 0 && K > 0, "vog_tc_gemm_lin2: bad dimension");
    if (M == 0) return 0;
    if (require_sm100("vog_tc_gemm_lin2")) return -1;
    VOG_REQUIRE(A && W, "vog_tc_gemm_lin2: null operand");
    VOG_REQUIRE(B > 0 && nfrm > 0 && nsrl > 0 && nppf2 > 0 && (long long)B * nfrm * nsrl * nppf2 == M,
                "vog_tc_gemm_lin2: M must be B*nfrm*nsrl*nppf2");
    VOG_REQUIRE(ncmp > 0 && nppf > 0 && nfrm0 > 0 && nfrm * nppf2 == ncmp * nfrm0 * nppf,
                "vog_tc_gemm_lin2: inconsistent frame/proposal grouping");
    TcEpilogue e;
    e.mode = 3; e.bias = bias; e.relu = 1; e.w2 = w2; e.b2 = b2;
    e.srl_msk = (const long long*)srl_msk; e.cmp_msk = (const long long*)cmp_msk;
    e.logits = logits; e.scores = scores;
    e.nfrm = nfrm; e.nsrl = nsrl; e.nppf2 = nppf2; e.ncmp = ncmp; e.nppf = nppf; e.nfrm0 = nfrm0; e.spat = spat;
    return tc_gemm(A, lda, W, ldw, M, N, K, tf32, N, e, nullptr, 0, (cudaStream_t)stream);
}

int vog_tc_gemm_gres(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int tf32,
                     int BN, const float* bias, int relu, const float* res_vis, int64_t ldv,
                     const float* res_lang, int64_t ldl, int dv, int nfrm, int nsrl, int nppf2,
                     float* out_f32, int64_t ldc, void* out_lp, int64_t ldlp, int lp_kind, void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0 && K > 0, "vog_tc_gemm_gres: bad dimension");
    if (M == 0 || N == 0) return 0;
    if (require_sm100("vog_tc_gemm_gres")) return -1;
    VOG_REQUIRE(A && W && res_vis && res_lang, "vog_tc_gemm_gres: null operand");
    VOG_REQUIRE(nfrm > 0 && nsrl > 0 && nppf2 > 0 && dv > 0 && dv < N && M % (nsrl * nppf2) == 0 &&
                (M / (nsrl * nppf2)) % nfrm == 0 && ldv >= dv && ldl >= N - dv,
                "vog_tc_gemm_gres: bad token geometry");
    TcEpilogue e;
    e.bias = bias; e.relu = relu; e.out_f32 = out_f32; e.ldc = ldc; e.out_lp = out_lp; e.ldlp = ldlp;
    e.lp_kind = lp_kind; e.rep = 1;
    e.res_vis = res_vis; e.ldv = ldv; e.res_lang = res_lang; e.ldl = ldl; e.dv = dv;
    e.nfrm = nfrm; e.nsrl = nsrl; e.nppf2 = nppf2;
    VOG_REQUIRE(!out_lp || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_tc_gemm_gres: bad lp_kind");
    return tc_gemm(A, lda, W, ldw, M, N, K, tf32, BN, e, nullptr, 0, (cudaStream_t)stream);
}

/* debug: device buffer of 8 int64 that receives clock64 stamps of CTA 0 of the next tc_gemm launches */
void vog_debug_gemm_trace(void* buf) { vog::tc_gemm_set_trace((long long*)buf); }

/* debug: device buffer of 8 int64 that receives per-phase cycle counts of one softmax warp */
void vog_debug_attn_prof(void* buf) { vog::tc_attn_set_prof((long long*)buf); }
/* A/B switch of the fused attention kernel: 1 = Q and P through shared memory, 2 = Q and P in tensor memory */
void vog_debug_attn_impl(int impl) { vog::tc_attn_set_impl(impl); }
/* force the thread-block-cluster size of the v2 attention kernel (1, 2, 4; 0 = automatic) */
void vog_debug_attn_cluster(int c) { vog::tc_attn_set_cluster(c); }

int64_t vog_tc_attn_workspace_bytes(int Bt, int N, int H)
{
    if (Bt <= 0 || N <= 0 || H <= 0) return 0;
    return tc_attn_workspace_bytes(Bt, N, H);
}

int vog_tc_attn_fwd(const void* q, const void* k, const void* v, int Bt, int N, int H, int dhp,
                    const int* dh, float inv_scale, int bias_mode, const float* a, int nbox,
                    const float* bpe, const float* dense, void* out, int64_t ldo, int out_kind,
                    void* workspace, int64_t workspace_bytes, void* stream)
{
    VOG_REQUIRE(Bt >= 0 && N >= 0, "vog_tc_attn_fwd: negative dimension");
    if (Bt == 0 || N == 0) return 0;
    if (require_sm100("vog_tc_attn_fwd")) return -1;
    VOG_REQUIRE(q && k && v && out && dh, "vog_tc_attn_fwd: null operand");
    VOG_REQUIRE(bias_mode >= 0 && bias_mode <= VOG_BIAS_RANK1_EXPANDED, "vog_tc_attn_fwd: bad bias_mode %d", bias_mode);
    return tc_attn(q, k, v, Bt, N, H, dhp, dh, inv_scale, bias_mode, a, nbox, bpe, dense, out, ldo,
                   out_kind, workspace, workspace_bytes, (cudaStream_t)stream);
}

int vog_tc_attn_fwd_train(const void* q, const void* k, const void* v, int Bt, int N, int H, int dhp, const int* dh,
                          float inv_scale, int bias_mode, const float* a, int nbox, const float* bpe, void* out,
                          int64_t ldo, int out_kind, void* workspace, int64_t workspace_bytes, float* lse, float drop_p,
                          uint64_t seed, void* stream)
{
    VOG_REQUIRE(Bt >= 0 && N >= 0, "vog_tc_attn_fwd_train: negative dimension");
    if (Bt == 0 || N == 0) return 0;
    if (require_sm100("vog_tc_attn_fwd_train")) return -1;
    VOG_REQUIRE(q && k && v && out && dh && lse, "vog_tc_attn_fwd_train: null operand");
    VOG_REQUIRE(bias_mode == 0 || bias_mode == 1, "vog_tc_attn_fwd_train: bias_mode %d (rank-1 or none)", bias_mode);
    return tc_attn(q, k, v, Bt, N, H, dhp, dh, inv_scale, bias_mode, a, nbox, bpe, nullptr, out, ldo, out_kind, workspace,
                   workspace_bytes, (cudaStream_t)stream, lse, drop_p, (unsigned long long)seed);
}

int64_t vog_tc_attn_bwd_workspace_bytes(int Bt, int N, int H)
{
    if (Bt <= 0 || N <= 0 || H <= 0) return 0;
    return tc_attn_bwd_workspace_bytes(Bt, N, H);
}

int vog_tc_attn_bwd(const void* q, const void* k, const void* v, const void* o, int64_t ldo, const void* dout,
                    int64_t lddo, const float* lse, int Bt, int N, int H, int dhp, const int* dh, float inv_scale,
                    int bias_mode, const float* a, int nbox, const float* bpe, void* dqkv, int64_t ldg, float* da,
                    float* dbpe, void* workspace, int64_t workspace_bytes, float drop_p, uint64_t seed, void* stream)
{
    VOG_REQUIRE(Bt >= 0 && N >= 0, "vog_tc_attn_bwd: negative dimension");
    if (Bt == 0 || N == 0) return 0;
    if (require_sm100("vog_tc_attn_bwd")) return -1;
    VOG_REQUIRE(q && k && v && o && dout && lse && dqkv && dh, "vog_tc_attn_bwd: null operand");
    return tc_attn_bwd(q, k, v, o, ldo, dout, lddo, lse, Bt, N, H, dhp, dh, inv_scale, bias_mode, a, nbox, bpe, dqkv, ldg,
                       da, dbpe, workspace, workspace_bytes, drop_p, (unsigned long long)seed, (cudaStream_t)stream);
}

int vog_pack_weights(const float* wq, const float* wk, const float* wv, const float* wo, int d, int H, const int* dh,
                     int dhp, int lp_kind, void* wqkv, void* wo_p, void* stream)
{
    VOG_REQUIRE(wq && wk && wv && wo && dh && wqkv && wo_p, "vog_pack_weights: null operand");
    return pack_weights(wq, wk, wv, wo, d, H, dh, dhp, lp_kind, wqkv, wo_p, (cudaStream_t)stream);
}

int64_t vog_workspace_bytes(int op, int a, int b, int c, int d, int e)
{
    switch (op) {
    case VOG_WS_TC_GEMM: return vog_tc_gemm_workspace_bytes(a, b, c, d, e);
    case VOG_WS_TC_ATTN: return vog_tc_attn_workspace_bytes(a, b, c);
    case VOG_WS_TC_ATTN_BWD: return vog_tc_attn_bwd_workspace_bytes(a, b, c);
    case VOG_WS_LSTM: return vog_lstm_workspace_bytes(a, b);
    case VOG_WS_LOSS: return vog_loss_workspace_bytes(a, b, c);
    case VOG_WS_LSTM_BWD: return vog_lstm_bwd_workspace_bytes(a, b);
    default: return -1;
    }
}

int vog_dropout(const float* x, int64_t ldx, const float* residual, int64_t ldr, float* out, int64_t ldo, void* out_lp,
                int64_t ldlp, int lp_kind, int64_t M, int N, float p, uint64_t seed, int stream_id, void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0, "vog_dropout: negative dimension");
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(x && (out || out_lp), "vog_dropout: null operand");
    return dropout_apply(x, ldx, residual, ldr, out, ldo, out_lp, ldlp, lp_kind, M, N, p, (unsigned long long)seed,
                         (unsigned int)stream_id, (cudaStream_t)stream);
}

int vog_tc_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, int K, int N1, int N2, float* C,
                   int64_t ldc, void* stream)
{
    VOG_REQUIRE(K >= 0 && N1 >= 0 && N2 >= 0, "vog_tc_gemm_tn: negative dimension");
    if (K == 0 || N1 == 0 || N2 == 0) return 0;
    if (require_sm100("vog_tc_gemm_tn")) return -1;
    VOG_REQUIRE(A && B && C, "vog_tc_gemm_tn: null operand");
    return tc_gemm_tn(A, lda, B, ldb, K, N1, N2, C, ldc, (cudaStream_t)stream);
}

int64_t vog_lstm_workspace_bytes(int Bq, int H) { return lstm_workspace_bytes(Bq, H); }

/* SM partitioning between concurrent branches of one forward (host-side state read at launch / graph-capture time):
 * vog_lstm_set_max_ctas(n > 0) runs the following recurrence launches on at most n SMs (weight-streaming kernel with
 * many hidden units per CTA); vog_set_reserved_sms(n) makes the persistent GEMMs size their grids to SMs - n.  0 resets. */
void vog_lstm_set_max_ctas(int n) { vog::lstm_set_max_ctas(n); }
void vog_set_reserved_sms(int n) { vog::tc_gemm_set_reserved_sms(n); }

/* debug / A-B testing: 1 = always use the weight-streaming recurrence kernel */
void vog_debug_lstm_force_streaming(int on) { vog::lstm_force_streaming(on); }
/* A/B switch of the h_t exchange of the weight-resident kernel: 4 = automatic (default), 2 = self-tagged per-CTA records, 3 = records + two hidden units per warp, 0 = tagged 64-bit words, 1 = per-CTA flags */
void vog_debug_lstm_exchange(int mode) { vog::lstm_set_exchange(mode); }
/* debug: device buffer of 8 int64: accumulated clock64 cycles of CTA 0 in {matvec, reduce, cell + publish,
 * poll, barrier} and the step count of the weight-resident recurrence kernel */
void vog_debug_lstm_trace(void* buf) { vog::lstm_set_trace((long long*)buf); }

int vog_lstm_layer_fwd(const float* gx, int64_t ldg, const float* whh, const int64_t* lens, int T, int Bq,
                       int H, void* out_lp, int64_t ld_out, int lp_kind, void* workspace, void* stream)
{
    VOG_REQUIRE(T >= 0 && Bq >= 0 && H > 0, "vog_lstm_layer_fwd: bad dimension");
    if (T == 0 || Bq == 0) return 0;
    VOG_REQUIRE(gx && whh && lens && out_lp && workspace, "vog_lstm_layer_fwd: null operand");
    return lstm_layer_fwd(gx, ldg, whh, (const long long*)lens, T, Bq, H, out_lp, ld_out, lp_kind, workspace,
                          (cudaStream_t)stream);
}

int vog_lstm_layer_fwd_train(const float* gx, int64_t ldg, const float* whh, const int64_t* lens, int T, int Bq,
                             int H, void* out_lp, int64_t ld_out, int lp_kind, void* workspace, float* acts,
                             void* stream)
{
    VOG_REQUIRE(T >= 0 && Bq >= 0 && H > 0, "vog_lstm_layer_fwd_train: bad dimension");
    if (T == 0 || Bq == 0) return 0;
    VOG_REQUIRE(gx && whh && lens && out_lp && workspace && acts, "vog_lstm_layer_fwd_train: null operand");
    return lstm_layer_fwd(gx, ldg, whh, (const long long*)lens, T, Bq, H, out_lp, ld_out, lp_kind, workspace,
                          (cudaStream_t)stream, acts);
}

int vog_build_xmul(const float* vis, const float* lang, float* out, void* out_lp, int lp_kind, int B,
                   int nfrm, int nsrl, int nppf2, int dv, int dl, void* stream)
{
    VOG_REQUIRE(B >= 0 && nfrm >= 0 && nsrl >= 0 && nppf2 >= 0, "vog_build_xmul: negative dimension");
    if ((long long)B * nfrm * nsrl * nppf2 == 0) return 0;
    VOG_REQUIRE(vis && lang && (out || out_lp), "vog_build_xmul: null operand");
    VOG_REQUIRE(!out_lp || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_build_xmul: bad lp_kind");
    return build_xmul(vis, lang, out, out_lp, lp_kind, B, nfrm, nsrl, nppf2, dv, dl, (cudaStream_t)stream);
}

int vog_lin2_tail(const float* h, int ldh, const float* w2, const float* b2, const int64_t* srl_msk,
                  const int64_t* cmp_msk, float* logits, float* scores, int B, int nfrm, int nsrl, int nppf2,
                  int K, int ncmp, int nppf, int nfrm0, int spat, void* stream)
{
    VOG_REQUIRE(B >= 0 && nfrm >= 0 && nsrl >= 0 && nppf2 >= 0, "vog_lin2_tail: negative dimension");
    if ((long long)B * nfrm * nsrl * nppf2 == 0) return 0;
    VOG_REQUIRE(h && w2 && b2 && srl_msk && cmp_msk && logits && scores, "vog_lin2_tail: null operand");
    VOG_REQUIRE(ncmp > 0 && nppf > 0 && nfrm0 > 0 && nfrm * nppf2 == ncmp * nfrm0 * nppf,
                "vog_lin2_tail: inconsistent frame/proposal grouping");
    return lin2_tail(h, ldh, w2, b2, (const long long*)srl_msk, (const long long*)cmp_msk, logits, scores, B, nfrm,
                     nsrl, nppf2, K, ncmp, nppf, nfrm0, spat, (cudaStream_t)stream);
}

int vog_lang_embed(const int64_t* words, int nwords, const int64_t* mask, int T, const float* emb, int E,
                   int64_t pad_idx, int Bq, void* out_lp, int lp_kind, void* stream)
{
    VOG_REQUIRE(T >= 0 && Bq >= 0 && nwords > 0 && E > 0, "vog_lang_embed: bad dimension");
    if (T * Bq == 0) return 0;
    VOG_REQUIRE(words && mask && emb && out_lp, "vog_lang_embed: null operand");
    VOG_REQUIRE(lp_kind == 0 || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_lang_embed: bad lp_kind");
    return lang_embed((const long long*)words, nwords, (const long long*)mask, T, emb, E, pad_idx, Bq, out_lp,
                      lp_kind, (cudaStream_t)stream);
}

int vog_lang_gather(const float* full, int D, const int64_t* cap, int T, int Bq, int nsrl, void* out_lp,
                    int lp_kind, void* stream)
{
    VOG_REQUIRE(T > 0 && Bq >= 0 && nsrl >= 0 && D > 0, "vog_lang_gather: bad dimension");
    if (Bq * nsrl == 0) return 0;
    VOG_REQUIRE(full && cap && out_lp, "vog_lang_gather: null operand");
    VOG_REQUIRE(lp_kind == 0 || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_lang_gather: bad lp_kind");
    return lang_gather(full, D, (const long long*)cap, T, Bq, nsrl, out_lp, lp_kind, (cudaStream_t)stream);
}

int vog_mask_rows(const float* x, const int64_t* msk, int rows, int D, float* out, void* out_lp, int lp_kind,
                  void* stream)
{
    VOG_REQUIRE(rows >= 0 && D > 0, "vog_mask_rows: bad dimension");
    if (rows == 0) return 0;
    VOG_REQUIRE(x && msk && out, "vog_mask_rows: null operand");
    VOG_REQUIRE(!out_lp || lp_kind == VOG_LP_BF16 || lp_kind == VOG_LP_TF32, "vog_mask_rows: bad lp_kind");
    return mask_rows(x, (const long long*)msk, rows, D, out, out_lp, lp_kind, (cudaStream_t)stream);
}

int vog_loss_bwd(const float* logits, const uint8_t* targets, const void* workspace, const float* grad_out, float* grad_logits,
                 int B, int nsrl, int P, void* stream)
{
    VOG_REQUIRE(logits && targets && workspace && grad_out && grad_logits, "vog_loss_bwd: null operand");
    return loss_bwd(logits, targets, workspace, grad_out, grad_logits, B, nsrl, P, (cudaStream_t)stream);
}

int64_t vog_loss_workspace_bytes(int B, int nsrl, int P)
{
    if (B <= 0 || nsrl <= 0 || P <= 0) return 0;
    return loss_workspace_bytes(B, nsrl, P);
}

int vog_loss_fwd(const float* logits, const float* props, int pdim, const float* gt, const uint8_t* frm_mask,
                 const uint8_t* pnt_mask, const int64_t* srl_boxes, const int64_t* srl_lens,
                 const int64_t* arg_boxes_mask, const int64_t* cmp_msk, const int64_t* target_cmp, int B, int nsrl,
                 int nb, int P, int K, int ncmp, int nppf, int spat, float loss_lambda, uint8_t* targets,
                 void* workspace, float* loss, void* stream)
{
    VOG_REQUIRE(B >= 0, "vog_loss_fwd: negative batch");
    VOG_REQUIRE(B > 0, "vog_loss_fwd: the mean over an empty batch is undefined");
    VOG_REQUIRE(logits && props && gt && frm_mask && pnt_mask && srl_boxes && srl_lens && arg_boxes_mask && cmp_msk &&
                target_cmp && workspace && loss, "vog_loss_fwd: null operand");
    VOG_REQUIRE(pdim >= 4, "vog_loss_fwd: proposals need >= 4 columns");
    return loss_fwd(logits, props, pdim, gt, frm_mask, pnt_mask, (const long long*)srl_boxes, (const long long*)srl_lens,
                    (const long long*)arg_boxes_mask, (const long long*)cmp_msk, (const long long*)target_cmp, B, nsrl,
                    nb, P, K, ncmp, nppf, spat, loss_lambda, targets, workspace, loss, (cudaStream_t)stream);
}

// ---- training step: backward entry points (train_f32.cu) ------------------------------------------------------
int vog_sgemm_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                      const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu, int accumulate,
                      void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0 && K >= 0, "vog_sgemm_strided: negative dimension");
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(A && B && C && ldc >= N, "vog_sgemm_strided: null operand / bad ldc");
    return sgemm_strided(A, sam, sak, B, sbk, sbn, bias, C, ldc, M, N, K, relu, accumulate, (cudaStream_t)stream);
}

int vog_colsum_acc(const float* x, int64_t ldx, float* out, int64_t M, int N, void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0, "vog_colsum_acc: negative dimension");
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(x && out && ldx >= N, "vog_colsum_acc: null operand / bad ldx");
    return colsum_acc(x, ldx, out, M, N, (cudaStream_t)stream);
}

int vog_relu_bwd(const float* dy, int64_t ldy, const void* act, int64_t lda, int act_kind, float* out, int64_t ldo,
                 void* out_lp, int64_t ldlp, int lp_kind, float* dbias, int64_t M, int N, void* stream)
{
    VOG_REQUIRE(M >= 0 && N >= 0, "vog_relu_bwd: negative dimension");
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(dy && act && (out || out_lp || dbias), "vog_relu_bwd: null operand");
    VOG_REQUIRE(!out_lp || (lp_kind >= 0 && lp_kind <= 2), "vog_relu_bwd: bad lp_kind");
    return relu_bwd(dy, ldy, act, lda, act_kind, out, ldo, out_lp, ldlp, lp_kind, dbias, M, N, (cudaStream_t)stream);
}

int vog_layernorm_bwd(const float* dy, int64_t ldy, const float* x, int64_t ldx, const float* gamma, float* dx,
                      int64_t lddx, void* dx_lp, int64_t ldlp, int lp_kind, float* dgamma, float* dbeta, float* dxsum,
                      int64_t M, int d, float eps, void* stream)
{
    VOG_REQUIRE(M >= 0 && d > 0, "vog_layernorm_bwd: bad dimension");
    if (M == 0) return 0;
    VOG_REQUIRE(dy && x && gamma && dgamma && dbeta, "vog_layernorm_bwd: null operand");
    VOG_REQUIRE(!dx_lp || (lp_kind >= 0 && lp_kind <= 2), "vog_layernorm_bwd: bad lp_kind");
    return layernorm_bwd(dy, ldy, x, ldx, gamma, dx, lddx, dx_lp, ldlp, lp_kind, dgamma, dbeta, dxsum, M, d, eps,
                         (cudaStream_t)stream);
}

int vog_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ld, const float* out, int64_t ldo,
                     const float* dout, int64_t lddo, const float* lse, float* delta, float* dq, float* dk, float* dv,
                     int64_t ldg, int Bt, int N, int H, const int* off, const int* dh, float inv_scale, int bias_mode,
                     const float* a, int nbox, const float* bpe, const float* dense, float* da, float* dbpe,
                     float* ddense, float drop_p, uint64_t seed, void* stream)
{
    VOG_REQUIRE(Bt >= 0 && N >= 0, "vog_attn_bwd_f32: negative dimension");
    if (Bt == 0 || N == 0) return 0;
    VOG_REQUIRE(q && k && v && out && dout && lse && delta && dq && dk && dv && off && dh, "vog_attn_bwd_f32: null operand");
    VOG_REQUIRE(bias_mode >= 0 && bias_mode <= 2, "vog_attn_bwd_f32: bad bias_mode %d", bias_mode);
    return attn_bwd_f32(q, k, v, ld, out, ldo, dout, lddo, lse, delta, dq, dk, dv, ldg, Bt, N, H, off, dh, inv_scale,
                        bias_mode, a, nbox, bpe, dense, da, dbpe, ddense, (cudaStream_t)stream, drop_p, (unsigned long long)seed);
}

int vog_pe_project_bwd(const float* props, int ldp, const float* da, float* dW, int rows, int H, float vid_w,
                       float vid_h, float fdiv, void* stream)
{
    VOG_REQUIRE(rows >= 0, "vog_pe_project_bwd: negative dimension");
    if (rows == 0) return 0;
    VOG_REQUIRE(props && da && dW && ldp >= 5, "vog_pe_project_bwd: null operand / bad ldp");
    return pe_project_bwd(props, ldp, da, dW, rows, H, vid_w, vid_h, fdiv, (cudaStream_t)stream);
}

int vog_xmul_bwd(const float* dtok, float* dvis, float* dlang, int B, int nfrm, int nsrl, int nppf2, int dv, int dl,
                 void* stream)
{
    VOG_REQUIRE(B >= 0 && nfrm >= 0 && nsrl >= 0 && nppf2 >= 0, "vog_xmul_bwd: negative dimension");
    if ((long long)B * nfrm * nsrl * nppf2 == 0) return 0;
    VOG_REQUIRE(dtok && dvis && dlang, "vog_xmul_bwd: null operand");
    return xmul_bwd(dtok, dvis, dlang, B, nfrm, nsrl, nppf2, dv, dl, (cudaStream_t)stream);
}

int vog_seg_rep_bwd(const float* dx, const float* x, int ld, int pe, int se, int nppf, float* dseg, int64_t nslots,
                    void* stream)
{
    VOG_REQUIRE(nslots >= 0 && nppf >= 1, "vog_seg_rep_bwd: bad dimension");
    if (nslots == 0) return 0;
    VOG_REQUIRE(dx && x && dseg && ld >= pe + se, "vog_seg_rep_bwd: null operand / bad ld");
    return seg_rep_bwd(dx, x, ld, pe, se, nppf, dseg, nslots, (cudaStream_t)stream);
}

int vog_lin2_bwd(const float* dlogits, const void* h, int64_t ldh, int h_kind, const float* w2, float* dh, void* dh_lp,
                 int lp_kind, float* dw2, float* db2, float* db1, int64_t M, int K, int nfrm, int nsrl, int nppf2,
                 void* stream)
{
    VOG_REQUIRE(M >= 0, "vog_lin2_bwd: negative dimension");
    if (M == 0) return 0;
    VOG_REQUIRE(dlogits && h && w2 && dw2 && db2 && db1 && (dh || dh_lp), "vog_lin2_bwd: null operand");
    VOG_REQUIRE(nfrm > 0 && nsrl > 0 && nppf2 > 0 && M % ((long long)nfrm * nsrl * nppf2) == 0, "vog_lin2_bwd: bad geometry");
    return lin2_bwd(dlogits, h, ldh, h_kind, w2, dh, dh_lp, lp_kind, dw2, db2, db1, M, K, nfrm, nsrl, nppf2,
                    (cudaStream_t)stream);
}

int vog_lang_gather_bwd(const float* dcat, int D, const int64_t* cap, int T, int Bq, int nsrl, float* dfull,
                        void* stream)
{
    if (Bq * nsrl == 0) return 0;
    VOG_REQUIRE(dcat && cap && dfull, "vog_lang_gather_bwd: null operand");
    return lang_gather_bwd(dcat, D, (const long long*)cap, T, Bq, nsrl, dfull, (cudaStream_t)stream);
}

int vog_lang_embed_bwd(const int64_t* words, int nwords, const int64_t* mask, int T, const float* dx, int E,
                       int64_t pad_idx, int Bq, const int64_t* lens, float* demb, void* stream)
{
    if (T * Bq == 0) return 0;
    VOG_REQUIRE(words && mask && dx && demb, "vog_lang_embed_bwd: null operand");
    return lang_embed_bwd((const long long*)words, nwords, (const long long*)mask, T, dx, E, pad_idx, Bq,
                          (const long long*)lens, demb, (cudaStream_t)stream);
}

int vog_lstm_hprev(const float* hout, const int64_t* lens, float* hprev, int T, int Bq, int H, void* stream)
{
    if (T * Bq == 0) return 0;
    VOG_REQUIRE(hout && lens && hprev && H > 0, "vog_lstm_hprev: null operand");
    return lstm_hprev(hout, (const long long*)lens, hprev, T, Bq, H, (cudaStream_t)stream);
}

int vog_lstm_scan(const float* G, const int64_t* lens, float* acts, int T, int Bq, int H, void* stream)
{
    if (T * Bq == 0) return 0;
    VOG_REQUIRE(G && lens && acts && H > 0, "vog_lstm_scan: null operand");
    return lstm_scan(G, (const long long*)lens, acts, T, Bq, H, (cudaStream_t)stream);
}

int64_t vog_lstm_bwd_workspace_bytes(int Bq, int H) { return lstm_bwd_workspace_bytes(Bq, H); }
/* A/B switch: 0 = T per-step launches even where the persistent weight-resident backward kernel applies */
void vog_debug_lstm_bwd_resident(int on) { vog::lstm_bwd_set_resident(on); }

int vog_lstm_bwd_steps(const float* dout, const float* acts, const float* whh_t, const float* whh, const int64_t* lens,
                       float* dG, void* workspace, int64_t workspace_bytes, int T, int Bq, int H, void* stream)
{
    if (T * Bq == 0) return 0;
    VOG_REQUIRE(dout && acts && whh_t && lens && dG && workspace && H > 0, "vog_lstm_bwd_steps: null operand");
    return lstm_bwd_steps(dout, acts, whh_t, whh, (const long long*)lens, dG, (float*)workspace, workspace_bytes, T, Bq, H,
                          (cudaStream_t)stream);
}

}  // extern "C"
