// Internal C++ launch API of libvog_b200 (one function per kernel family).  The exported C ABI in
// vog_abi.cu (declared in include/vog_b200.h) is a thin argument-checking shim over these.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#define VOG_MAX_HEADS 8

namespace vog {

// ---- fp32_path.cu : exact-fp32 CUDA-core kernels --------------------------------------------
int sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias,
             const float* R, int ldr, float* C, int ldc, int M, int N, int K, int relu,
             cudaStream_t st);
int attn_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo,
             int Bt, int N, int H, const int* off, const int* dh, float inv_scale,
             int bias_mode, const float* a, int nbox, const float* bpe, const float* dense,
             cudaStream_t st, float* lse = nullptr, float drop_p = 0.f, unsigned long long seed = 0);
int add_layernorm(const float* x, int ldx, const float* r, int ldr, const float* w, const float* b,
                  float* out, int ldo, void* out_lp, int ldlp, int lp_kind, int M, int d, float eps,
                  cudaStream_t st);
int pe_project(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw,
               float vh, float fdiv, float scale, cudaStream_t st);
int pe_project_expand(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw, float vh,
                      float fdiv, float scale, float* ak, int Bt, int N, int nbox, int ld, float c, cudaStream_t st);
int tc_attn_key_ld(int N);            // row length of the expanded key factors (N rounded up to the key tile)
float tc_attn_key_scale(float inv_scale);   // the factor the attention kernel applies to them (exp2 domain)
int select_fwd(const float* scores, const float* props, int pdim, float* boxes, float* out_scores,
               long long* indexs, int B, int nsrl, int ncmp, int nfrm, int nppf, int spat,
               cudaStream_t st, const float* fin = nullptr);
int concat_videos(const float* feat, int D, const float* seg, int Ds, const float* props, int pdim, float* feat_out,
                  float* seg_out, float* props_out, int B, int ncmp, int nfrm, int nppf, int spat, float shift,
                  cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
              long long step, double grad_scale, cudaStream_t st);
int verb_loss_fwd(const float* vidf, const long long* verb_cmp, const long long* vcc, int n, int m, float lambda,
                  float* loss, cudaStream_t st);
int sep_fin_scores(const float* logits, const float* vidf, const long long* srl_msk, const long long* verb_ind,
                   const long long* cmp_msk, float* fin_loss, float* fin_eval, int Bq, int nsrl, int P1, cudaStream_t st);

// ---- tc_gemm.cu : tcgen05 / TMA / TMEM GEMM ---------------------------------------------------
struct TcEpilogue {
    int mode = 0;                 // 0 standard, 1 QKV scatter
    const float* bias = nullptr;  // [N]
    int relu = 0;
    const float* residual = nullptr; long long ldr = 0;   // fp32 [M,N]
    float* out_f32 = nullptr; long long ldc = 0;
    void* out_lp = nullptr; long long ldlp = 0; int lp_kind = 0;   // 1 bf16, 2 tf32-rounded fp32
    int rep = 1;                  // every output row m is written to rows m*rep .. m*rep+rep-1
    // mode 1: column block n_blk = which*H + h (BN == dhp); rows m = bt*seq_n + i
    __nv_bfloat16* q = nullptr; __nv_bfloat16* k = nullptr;   // [Bt,H,seq_n,dhp]
    __nv_bfloat16* v = nullptr;                               // [Bt,H,seq_n,dhp] (natural layout, like K)
    int seq_n = 0, n_heads = 0, dhp = 0;
    // mode 2: FACTORISED QKV projection of the multimodal transformer.  A token (bt, s, p) of sequence
    // bt = b*nfrm + f is [vis[bt*nppf2 + p] | lang[b*nsrl + s]], so W.token = W[:, :dv].vis + W[:, dv:].lang:
    // the GEMM runs over the VISUAL rows only (m = bt*nppf2 + p) and its epilogue writes every row
    // nsrl times - to tokens (bt, s, p), s < nsrl, seq_n = nsrl*nppf2 - adding lq[b*nsrl + s, col], the
    // separately projected language rows.  5x fewer projection FLOPs, x_mul never materialised.
    const float* lq = nullptr; long long ldq = 0;
    int nsrl = 0, nppf2 = 0, nfrm = 0;
    // mode 0, gathered residual: the residual row of token m = (bt, s, p) is
    // [res_vis[bt*nppf2 + p, 0:dv] | res_lang[b*nsrl + s, 0:N-dv]]  (dv % BN == 0)
    const float* res_vis = nullptr; long long ldv = 0;
    const float* res_lang = nullptr; long long ldl = 0;
    int dv = 0;
    // mode 3: scorer tail fused into the lin2[0] GEMM (BN == N <= 256, one column tile per row block):
    //   logit[m] = relu(acc[m,:] + bias) . w2 + b2, inverse regroup of token m = (b, f, s, p) to [B,nsrl,P],
    //   score = sigmoid(logit) * srl_msk[b,s] * cmp_msk[b, vid(p)]  -  the [M, N] hidden matrix is never written
    const float* w2 = nullptr; const float* b2 = nullptr;
    const long long* srl_msk = nullptr; const long long* cmp_msk = nullptr;
    float* logits = nullptr; float* scores = nullptr;
    int ncmp = 0, nppf = 0, nfrm0 = 0, spat = 0;
};
int tc_gemm(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int tf32,
            int BN, const TcEpilogue& epi, void* workspace, long long workspace_bytes, cudaStream_t st);
long long tc_gemm_workspace_bytes(int M, int N, int K, int tf32, int BN);
int num_sms();
void tc_gemm_set_reserved_sms(int n);    // persistent GEMM grids leave n SMs to a concurrent branch
void tc_gemm_set_trace(long long* buf);  // debug: 8 clock64 stamps of CTA 0
// ---- tc_attn.cu : fused relative-position-bias attention (tcgen05) ---------------------------
int tc_attn(const void* q, const void* k, const void* v, int Bt, int N, int H, int dhp,
            const int* dh, float inv_scale, int bias_mode, const float* a, int nbox, const float* bpe,
            const float* dense, void* out, long long ldo, int out_kind, void* workspace,
            long long workspace_bytes, cudaStream_t st, float* lse = nullptr, float drop_p = 0.f,
            unsigned long long seed = 0);
long long tc_attn_workspace_bytes(int Bt, int N, int H);
// ---- tc_attn_bwd.cu : attention backward on the tensor cores ------------------------------------
long long tc_attn_bwd_workspace_bytes(int Bt, int N, int H);
int tc_attn_bwd(const void* q, const void* k, const void* v, const void* o, long long ldo, const void* dout, long long lddo,
                const float* lse, int Bt, int N, int H, int dhp, const int* dh, float inv_scale, int bias_mode,
                const float* a, int nbox, const float* bpe, void* dqkv, long long ldg, float* da, float* dbpe,
                void* workspace, long long workspace_bytes, float drop_p, unsigned long long seed, cudaStream_t st);
// ---- tc_gemm_tn.cu : C[N1,N2] += A[K,N1]^T . B[K,N2] (weight gradients) --------------------------
int tc_gemm_tn(const void* A, long long lda, const void* B, long long ldb, int K, int N1, int N2, float* C,
               long long ldc, cudaStream_t st);
void tc_attn_set_impl(int impl);         // debug: 1 = Q/P via shared memory, 2 = Q/P in tensor memory (default)
void tc_attn_set_cluster(int c);         // debug: force the v2 cluster size (0 = auto)
void tc_attn_set_prof(long long* buf);   // debug: phase cycle counters of one softmax warp
int cast_lp(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols, int kind,
            cudaStream_t st);

// ---- lstm_rec.cu : persistent bidirectional LSTM recurrence -----------------------------------
long long lstm_workspace_bytes(int Bq, int H);
void lstm_set_trace(long long* buf);     // debug: phase cycle counters of CTA 0
long long* lstm_get_trace();
void lstm_set_exchange(int mode);        // debug: 4 automatic (default), 2 self-tagged records, 3 records + two units per warp, 0 tagged words, 1 flags
void lstm_set_max_ctas(int n);           // > 0: run the recurrence on at most n SMs (weight-streaming kernel)
void lstm_force_streaming(int on);       // debug: disable the weight-resident kernel
int lstm_layer_fwd(const float* gx, long long ldg, const float* whh, const long long* lens, int T, int Bq,
                   int H, void* out_lp, long long ld_out, int lp_kind, void* workspace, cudaStream_t st,
                   float* acts = nullptr);

// ---- fused_glue.cu ------------------------------------------------------------------------------
int build_xmul(const float* vis, const float* lang, float* out, void* out_lp, int lp_kind, int B, int nfrm,
               int nsrl, int nppf2, int dv, int dl, cudaStream_t st);
int lin2_tail(const float* h, int ldh, const float* w2, const float* b2, const long long* srl_msk,
              const long long* cmp_msk, float* logits, float* scores, int B, int nfrm, int nsrl, int nppf2,
              int K, int ncmp, int nppf, int nfrm0, int spat, cudaStream_t st);

int pack_weights(const float* wq, const float* wk, const float* wv, const float* wo, int d, int H, const int* dh, int dhp,
                 int lp_kind, void* wqkv, void* wo_p, cudaStream_t st);
int lang_embed(const long long* words, int nwords, const long long* mask, int T, const float* emb, int E,
               long long pad_idx, int Bq, void* out_lp, int lp_kind, cudaStream_t st);
int lang_gather(const float* full, int D, const long long* cap, int T, int Bq, int nsrl, void* out_lp, int lp_kind,
                cudaStream_t st);
int mask_rows(const float* x, const long long* msk, int rows, int D, float* out, void* out_lp, int lp_kind,
              cudaStream_t st);

// ---- train_f32.cu : backward kernels (exact fp32 + element-wise kernels shared by every compute mode) ----------
int sgemm_strided(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                  const float* bias, float* C, long long ldc, int M, int N, int K, int relu, int accumulate,
                  cudaStream_t st);
int colsum_acc(const float* x, long long ldx, float* out, long long M, int N, cudaStream_t st);
int relu_bwd(const float* dy, long long ldy, const void* act, long long lda, int act_kind, float* out, long long ldo,
             void* out_lp, long long ldlp, int lp_kind, float* dbias, long long M, int N, cudaStream_t st);
int layernorm_bwd(const float* dy, long long ldy, const float* x, long long ldx, const float* gamma, float* dx,
                  long long lddx, void* dx_lp, long long ldlp, int lp_kind, float* dgamma, float* dbeta, float* dxsum,
                  long long M, int d, float eps, cudaStream_t st);
int attn_bwd_f32(const float* q, const float* k, const float* v, long long ld, const float* out, long long ldo,
                 const float* dout, long long lddo, const float* lse, float* delta, float* dq, float* dk, float* dv,
                 long long ldg, int Bt, int N, int H, const int* off, const int* dh, float inv_scale, int bias_mode,
                 const float* a, int nbox, const float* bpe, const float* dense, float* da, float* dbpe, float* ddense,
                 cudaStream_t st, float drop_p = 0.f, unsigned long long seed = 0);
int dropout_apply(const float* x, long long ldx, const float* res, long long ldr, float* out, long long ldo, void* out_lp,
                  long long ldlp, int lp_kind, long long M, int N, float p, unsigned long long seed, unsigned int stream,
                  cudaStream_t st);
int pe_project_bwd(const float* props, int ldp, const float* da, float* dW, int rows, int H, float vw, float vh,
                   float fdiv, cudaStream_t st);
int xmul_bwd(const float* dtok, float* dvis, float* dlang, int B, int nfrm, int nsrl, int nppf2, int dv, int dl,
             cudaStream_t st);
int seg_rep_bwd(const float* dx, const float* x, int ld, int pe, int se, int nppf, float* dseg, long long nslots,
                cudaStream_t st);
int lin2_bwd(const float* dlogits, const void* h, long long ldh, int h_kind, const float* w2, float* dh, void* dh_lp,
             int lp_kind, float* dw2, float* db2, float* db1, long long M, int K, int nfrm, int nsrl, int nppf2,
             cudaStream_t st);
int lang_gather_bwd(const float* dcat, int D, const long long* cap, int T, int Bq, int nsrl, float* dfull, cudaStream_t st);
int lang_embed_bwd(const long long* words, int nwords, const long long* mask, int T, const float* dx, int E,
                   long long pad_idx, int Bq, const long long* lens, float* demb, cudaStream_t st);
int lstm_hprev(const float* hout, const long long* lens, float* hprev, int T, int Bq, int H, cudaStream_t st);
int lstm_scan(const float* G, const long long* lens, float* acts, int T, int Bq, int H, cudaStream_t st);
int lstm_bwd_steps(const float* dout, const float* acts, const float* whh_t, const float* whh, const long long* lens, float* dG,
                   float* carry_ws, long long ws_bytes, int T, int Bq, int H, cudaStream_t st);
// ---- lstm_bwd.cu : weight-resident persistent backward through time -----------------------------
long long lstm_bwd_workspace_bytes(int Bq, int H);
void lstm_bwd_set_resident(int on);     // debug: 0 = per-step kernels even where the persistent kernel applies
int lstm_bwd_resident(const float* dout, const float* acts, const float* whh_t, const float* whh, const long long* lens, float* dG,
                      void* ws, long long ws_bytes, int T, int Bq, int H, cudaStream_t st);   // 1 launched, 0 n/a, -1 error

// ---- loss_fwd.cu : grounding loss, forward ----------------------------------------------------
long long loss_workspace_bytes(int B, int nsrl, int P);
int loss_bwd(const float* logits, const unsigned char* targets, const void* workspace, const float* grad_out, float* grad,
             int B, int nsrl, int P, cudaStream_t st);
int loss_fwd(const float* logits, const float* props, int pdim, const float* gt, const unsigned char* frm_mask,
             const unsigned char* pnt_mask, const long long* srl_boxes, const long long* srl_lens,
             const long long* arg_boxes_mask, const long long* cmp_msk, const long long* target_cmp, int B,
             int nsrl, int nb, int P, int K, int ncmp, int nppf, int spat, float lambda, unsigned char* targets,
             void* workspace, float* loss, cudaStream_t st);

}  // namespace vog
