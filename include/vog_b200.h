/* libvog_b200 - C ABI of the B200-native VOGNet forward fusion path.
 *
 * The reference (TheShadow29/vognet-pytorch) is pure Python over PyTorch: it has no FFI layer, so
 * each entry point below replaces a *library-call sequence* of the reference, cited as
 * path:line relative to the reference root.  The Python host side
 * (vognet_pytorch_b200/_lib.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller, row-major, with explicit leading
 *     dimensions (in elements); nothing is allocated inside the library;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued, nothing synchronises;
 *   - return 0 on success, -1 on error; vog_last_error() returns the (thread-local) message;
 *   - thread-compatible: no mutable global state besides per-function CUDA attributes.
 */
#ifndef VOG_B200_H
#define VOG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VOG_MAX_HEADS 8

/* bias modes of the attention entry points */
#define VOG_BIAS_NONE  0   /* plain Attention          code/transformer_code.py:41-50            */
#define VOG_BIAS_RANK1 1   /* relu(a_i - a_j + b_h)    code/mdl_vog.py:456-490 in rank-1 form    */
#define VOG_BIAS_DENSE 2   /* dense x_pe [Bt,N,N,H]    code/transformer_code.py:271-273 operator */
#define VOG_BIAS_RANK1_EXPANDED 3   /* VOG_BIAS_RANK1 for vog_tc_attn_fwd when the workspace already holds the key factors
                                       (vog_pe_project_expand): no expansion pre-kernel */

/* low-precision copy kinds */
#define VOG_LP_NONE 0
#define VOG_LP_BF16 1
#define VOG_LP_TF32 2      /* fp32 container, rounded to nearest tf32 */

const char* vog_last_error(void);
int vog_abi_version(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py gpu_launches) */
long long vog_launch_count(void);
/* 1 when the current device is sm_100 (B200); the tcgen05 entry points refuse to run otherwise */
int vog_device_is_sm100(void);

/* ---------------------------------------------------------------------------------------------
 * exact-fp32 path (CUDA cores)
 * ------------------------------------------------------------------------------------------- */

/* C[M,N] = (relu?)(A[M,K] . W[N,K]^T + bias[N]) + R[M,N]        (bias, R nullable)
 * replaces nn.Linear (+ReLU) call sites: code/transformer_code.py:57-60,80-81,169-172,180,186;
 * code/mdl_vog.py:202-207,224-230,291-314,675-677. */
int vog_sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias,
                 const float* R, int ldr, float* C, int ldc, int M, int N, int K, int relu,
                 void* stream);

/* Multi-head attention with relative-position bias, all heads in one launch:
 *   out[bt*N+i, off[h]+c] = sum_j softmax_j((q_i.k_j + bias_h(i,j)) * inv_scale) v_j[c]
 * q,k,v: [Bt*N, ld] fp32, head h = columns off[h]..off[h]+dh[h] (torch.chunk split, uneven ok).
 * bias_mode RANK1: a [Bt*nbox, H] (vog_pe_project output, scale=1), bpe[H] (device); token t
 * uses box t % nbox.  bias_mode DENSE: dense [Bt,N,N,H].
 * replaces RelMultiHead/RelAttention + MultiHead/Attention: code/transformer_code.py:34-70,128-186
 * and the bias materialisation code/mdl_vog.py:477-488, utils/mdl_srl_utils.py:30-69. */
int vog_attn_fwd_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo,
                     int Bt, int N, int H, const int* off, const int* dh, float inv_scale,
                     int bias_mode, const float* a, int nbox, const float* bpe,
                     const float* dense, float* lse, float drop_p, uint64_t seed, void* stream);
/* lse (nullable): [Bt,H,N] log-sum-exp of the scaled scores, kept by the training forward for vog_attn_bwd_f32.
 * drop_p > 0: dropout on the probabilities after the softmax (code/transformer_code.py:153) with the same
 * counter-based mask (seed, sequence*H + head, query, key) as vog_tc_attn_fwd_train. */

/* out = LayerNorm(x + r) * w + b (eps inside the sqrt), r nullable; optional low-precision copy
 * out_lp (VOG_LP_*) for the next GEMM.  replaces ResidualBlock: code/transformer_code.py:21-31. */
int vog_add_layernorm(const float* x, int ldx, const float* r, int ldr, const float* w,
                      const float* b, float* out, int ldo, void* out_lp, int ldlp, int lp_kind,
                      int M, int d, float eps, void* stream);

/* a[row,h] = scale * W[h,:] . (x1/vw, y1/vh, x2/vw, y2/vh, frame/fdiv); props row stride ldp.
 * The rank-1 factor of Linear(5,H) applied to p_i - p_j: code/mdl_vog.py:446-451,459-463,480. */
int vog_pe_project(const float* props, int ldp, const float* W, float* a, int rows, int H,
                   float vw, float vh, float fdiv, float scale, void* stream);
/* vog_pe_project and, in the same launch, the per-key factors of the rank-1 bias that vog_tc_attn_fwd otherwise
 * expands in a pre-kernel of its own: key_factors (vog_tc_attn_workspace_bytes(Bt, N, H) bytes) for the attention
 * call with the same Bt, N, H, nbox and inv_scale, passed to it as its workspace with bias_mode
 * VOG_BIAS_RANK1_EXPANDED.  rows == Bt * nbox. */
int vog_pe_project_expand(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw, float vh,
                          float fdiv, float scale, int Bt, int N, int nbox, float inv_scale, void* key_factors,
                          int64_t key_factors_bytes, void* stream);

/* Selection: scores [B,nsrl,P] (= mdl_outs_eval), props [B,P,pdim] ->
 *   boxes [B,nsrl,ncmp,nfrm,pdim], out_scores [B,nsrl,ncmp,nfrm], indexs [B,nsrl,nfrm] (int64;
 *   argmax over vids for spat, zeros for temp).  Lowest index wins ties (torch.max semantics).
 * replaces EvaluatorSPAT/TEMP.get_out_results_boxes: code/eval_vsrl_corr.py:289-345,357-424. */
int vog_select_fwd(const float* scores, const float* props, int pdim, float* boxes,
                   float* out_scores, int64_t* indexs, int B, int nsrl, int ncmp, int nfrm,
                   int nppf, int spat, void* stream);

/* SEP selection: scores [B,ncmp,nsrl,nfrm*nppf] (= mdl_outs_eval of the SEP forward), props [B,ncmp,nfrm*nppf,pdim],
 * fin_scores [B,ncmp] -> boxes [B,nsrl,ncmp,nfrm,pdim], out_scores [B,nsrl,ncmp,nfrm], indexs [B,nsrl,nfrm] =
 * argmax over videos of fin_scores (lowest index wins ties).
 * replaces EvaluatorSEP.get_out_results_boxes: code/eval_vsrl_corr.py:162-220. */
int vog_select_sep_fwd(const float* scores, const float* props, int pdim, const float* fin_scores, float* boxes,
                       float* out_scores, int64_t* indexs, int B, int nsrl, int ncmp, int nfrm, int nppf, void* stream);

/* SEP fused per-video score: logits [Bq,nsrl,P1] (Bq = B*ncmp), vidf [Bq] (video-level verb logit), srl_msk [Bq,nsrl],
 * verb_ind [Bq], cmp_msk [Bq] (int64) -> fin_loss [Bq,nsrl] = max_p sigmoid(logits) with the verb slot replaced by
 * sigmoid(vidf), times srl_msk and cmp_msk; fin_eval [Bq] = sum_s(.. * srl_msk) / sum_s srl_msk * cmp_msk.
 * replaces ConcSEP.compute_fin_scores (use_vis_msk): code/mdl_conc_sep.py:62-117. */
int vog_sep_fin_scores(const float* logits, const float* vidf, const int64_t* srl_msk, const int64_t* verb_ind,
                       const int64_t* cmp_msk, float* fin_loss, float* fin_eval, int Bq, int nsrl, int P1, void* stream);

/* ---------------------------------------------------------------------------------------------
 * tensor-core path (tcgen05.mma + TMA + TMEM, sm_100a only; these refuse to run elsewhere)
 * ------------------------------------------------------------------------------------------- */

/* dst = bf16(src) (VOG_LP_BF16) or tf32-rounded fp32 (VOG_LP_TF32): A operand of the GEMMs below. */
int vog_cast_lp(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int cols,
                int kind, void* stream);

/* C = (relu?)(A[M,K] . W[N,K]^T + bias) + residual, A/W bf16 (tf32=0) or tf32-rounded fp32
 * (tf32=1), fp32 accumulation in TMEM.  128 x BN tiles (BN multiple of 32, <= 256), persistent
 * grid.  Outputs: out_f32 and/or a low-precision copy out_lp (lp_kind); every output row m is
 * written to rows m*rep .. m*rep+rep-1 (rep > 1 broadcasts a segment feature over the proposals of
 * its frame: code/mdl_conc_single.py:50-66,156-174).  Small-M problems split K across CTAs when
 * the caller passes a workspace of vog_tc_gemm_workspace_bytes() bytes (fp32 partials, reduced by
 * a second kernel); workspace may be NULL.  Same call sites as vog_sgemm_nt. */
int64_t vog_tc_gemm_workspace_bytes(int M, int N, int K, int tf32, int BN);
int vog_tc_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                int tf32, int BN, const float* bias, int relu, const float* residual, int64_t ldr,
                float* out_f32, int64_t ldc, void* out_lp, int64_t ldlp, int lp_kind, int rep,
                void* workspace, int64_t workspace_bytes, void* stream);

/* Fused Q/K/V projection: A[M,K] . Wqkv[3*H*dhp, K]^T with Wqkv = per-head zero-padded rows of
 * wq|wk|wv (dhp = head dim rounded up to 64).  Writes bf16 Q, K and V as [Bt,H,seq_n,dhp]
 * (M = Bt*seq_n), the layouts vog_tc_attn_fwd consumes (V as an MN-major tcgen05 operand).
 * replaces wq/wk/wv + chunk: code/transformer_code.py:180-183. */
int vog_tc_gemm_qkv(const void* A, int64_t lda, const void* Wqkv, int64_t ldw, int M, int K,
                    int tf32, int n_heads, int dhp, int seq_n, void* q, void* k, void* v, void* stream);

/* Fused attention, all heads and sequences in one launch (tcgen05 QK^T and PV, online softmax with
 * on-the-fly relative-position bias; the N x N matrices never reach HBM):
 *   out[bt*N+i, h*dhp+c] = sum_j softmax_j((q_i.k_j + bias_h(i,j)) * inv_scale) v_j[c]
 * q,k,v [Bt,H,N,dhp] bf16 (vog_tc_gemm_qkv layouts), dh[H] true head dims
 * (host array), bias as in vog_attn_fwd_f32 (a, bpe, dense are device pointers).  out is
 * [Bt*N, ldo >= H*dhp], bf16 (out_kind VOG_LP_BF16) or tf32-rounded fp32 (VOG_LP_TF32); padded
 * head columns are written as zeros.  VOG_BIAS_RANK1 needs a workspace of
 * vog_tc_attn_workspace_bytes() bytes (per-key bias factors expanded to [Bt*H, N] by a small
 * pre-kernel).  Same reference lines as vog_attn_fwd_f32. */
int64_t vog_tc_attn_workspace_bytes(int Bt, int N, int H);
int vog_tc_attn_fwd(const void* q, const void* k, const void* v, int Bt, int N, int H, int dhp,
                    const int* dh, float inv_scale, int bias_mode, const float* a, int nbox,
                    const float* bpe, const float* dense, void* out, int64_t ldo, int out_kind,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* FACTORISED Q/K/V projection of the multimodal transformer's first layer.  A token (bt, s, p) of the
 * per-frame sequence bt = b*nfrm + f is [vis[bt*nppf2 + p] | lang[b*nsrl + s]] (code/mdl_vog.py:316-344,
 * 693-699), so W.token = W[:, :dv].vis + W[:, dv:].lang: A [M = Bt*nppf2, K = dv] holds the VISUAL rows
 * only, Wvis = the first dv columns of the packed Wq|Wk|Wv (row stride ldw), lq [B*nsrl, ldq >= 3*H*dhp]
 * fp32 the separately projected language rows (a [B*nsrl x dl] GEMM through vog_tc_gemm).  The epilogue
 * writes every accumulator row nsrl times, adding lq[b*nsrl + s]: q,k,v [Bt,H,nsrl*nppf2,dhp] bf16
 * exactly as vog_tc_gemm_qkv would from the materialised [vis|lang] matrix - with
 * nsrl x fewer MMA FLOPs and without that matrix.  replaces concate_vis_lang_feats + regroup + wq/wk/wv:
 * code/mdl_vog.py:316-344,693-699, code/transformer_code.py:180. */
int vog_tc_gemm_qkv_factored(const void* A, int64_t lda, const void* Wvis, int64_t ldw, int M, int K,
                             int tf32, int n_heads, int dhp, const float* lq, int64_t ldq, int nfrm,
                             int nsrl, int nppf2, void* q, void* k, void* v, void* stream);

/* vog_tc_gemm with a GATHERED fp32 residual: the residual row of token m = (bt, s, p) is
 * [res_vis[bt*nppf2 + p, 0:dv] | res_lang[(bt / nfrm)*nsrl + s, 0:N-dv]] - the multimodal transformer's
 * input read from its two factors (dv % BN == 0, rows 16-byte aligned).  replaces the residual add of
 * the first RelEncoderLayer on the materialised concat: code/transformer_code.py:30-31,201-203. */
int vog_tc_gemm_gres(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int tf32,
                     int BN, const float* bias, int relu, const float* res_vis, int64_t ldv,
                     const float* res_lang, int64_t ldl, int dv, int nfrm, int nsrl, int nppf2,
                     float* out_f32, int64_t ldc, void* out_lp, int64_t ldlp, int lp_kind, void* stream);

/* The scorer: lin2[0] GEMM with lin2[2], the inverse regroup and the output masks fused into its epilogue
 * (N <= 256 = one column tile, so a thread owns a whole output row in tensor memory):
 *   logit = relu(A[m,:] . W^T + bias) . w2 + b2 for token m = ((b*nfrm + f)*nsrl + s)*nppf2 + p, written to
 *   logits / scores [B,nsrl,P] (P = nfrm*nppf2 = ncmp*nfrm0*nppf) with scores = sigmoid(logit) * srl_msk[b,s] *
 *   cmp_msk[b, vid(p)] (int64 masks).  The [M, N] hidden matrix never reaches HBM.  replaces lin2 + un-regroup +
 * masks: code/mdl_vog.py:224-230,675-677,724-737, code/mdl_conc_single.py:39-48,118-122,144-154. */
int vog_tc_gemm_lin2(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int tf32,
                     const float* bias, const float* w2, const float* b2, const int64_t* srl_msk,
                     const int64_t* cmp_msk, float* logits, float* scores, int B, int nfrm, int nsrl, int nppf2,
                     int ncmp, int nppf, int nfrm0, int spat, void* stream);

/* One layer of the bidirectional LSTM recurrence of the language encoder, both directions, all
 * timesteps, in one persistent launch (packed-sequence semantics from the device-side `lens`, no
 * host synchronisation):  gx [T*Bq, ldg >= 8H] = W_ih x + b_ih + b_hh for every (t, b) (time-major
 * rows t*Bq+b; forward gates in columns [0,4H), reverse in [4H,8H), gate order i,f,g,o), whh
 * [2,4H,H] fp32, lens [Bq] int64.  Writes h as [T*Bq, ld_out >= 2H] (forward | reverse) in bf16 or
 * tf32-rounded fp32, zeros where t >= len.  Any Bq (groups of 8 sequences per launch).  workspace:
 * vog_lstm_workspace_bytes() bytes.  replaces nn.LSTM + pack/pad:
 * utils/mdl_srl_utils.py:102-108,135-152. */
int64_t vog_lstm_workspace_bytes(int Bq, int H);
void vog_debug_lstm_force_streaming(int on);   /* debug: 1 = weight-streaming kernel even when H == 1024 */
int vog_lstm_layer_fwd(const float* gx, int64_t ldg, const float* whh, const int64_t* lens, int T,
                       int Bq, int H, void* out_lp, int64_t ld_out, int lp_kind, void* workspace,
                       void* stream);

/* Training variant: additionally keeps, for every live (t, b), the activations the step used - acts
 * [T*Bq, 2 dir, 6, H] = i, f, g, o, tanh(c_t), c_{t-1} - which is all vog_lstm_bwd_steps needs (no recompute). */
int vog_lstm_layer_fwd_train(const float* gx, int64_t ldg, const float* whh, const int64_t* lens, int T, int Bq,
                             int H, void* out_lp, int64_t ld_out, int lp_kind, void* workspace, float* acts,
                             void* stream);

/* SM partitioning between the concurrent branches of one forward (host-side state, read at launch / graph-capture
 * time).  vog_lstm_set_max_ctas(n > 0): the following vog_lstm_layer_fwd launches use at most n SMs (weight-streaming
 * kernel, many hidden units per CTA) instead of the weight-resident kernel that needs every SM; vog_set_reserved_sms(n):
 * the persistent vog_tc_gemm* grids leave n SMs free.  Used when the visual branch is much longer than the language
 * branch (100 proposals per frame): the recurrence then hides behind it instead of serialising with it.  0 resets. */
void vog_lstm_set_max_ctas(int n);
void vog_set_reserved_sms(int n);

/* Multimodal-transformer input: row ((b*nfrm+f)*nsrl+s)*nppf2+p of out / out_lp =
 * [vis[b*nfrm*nppf2 + f*nppf2 + p, 0:dv] | lang[b*nsrl + s, 0:dl]] as fp32 and/or low precision.
 * replaces concate_vis_lang_feats + the per-frame regroup: code/mdl_vog.py:316-344,693-699. */
int vog_build_xmul(const float* vis, const float* lang, float* out, void* out_lp, int lp_kind, int B,
                   int nfrm, int nsrl, int nppf2, int dv, int dl, void* stream);

/* Scorer tail: logit = h[row,0:K] . w2 + b2, inverse regroup to logits[B,nsrl,P] (P = nfrm*nppf2 =
 * ncmp*nfrm0*nppf), scores = sigmoid(logit) * srl_msk[b,s] * cmp_msk[b, vid(p)] (int64 masks).
 * replaces lin2[2], the un-regroup and the output masks: code/mdl_vog.py:675-677,724-737,
 * code/mdl_conc_single.py:39-48,118-122,144-154. */
int vog_lin2_tail(const float* h, int ldh, const float* w2, const float* b2, const int64_t* srl_msk,
                  const int64_t* cmp_msk, float* logits, float* scores, int B, int nfrm, int nsrl,
                  int nppf2, int K, int ncmp, int nppf, int nfrm0, int spat, void* stream);

/* Language-side glue.  vog_lang_embed: time-major LSTM input rows x_lp[t*Bq + b] = lp(emb[tok]) with
 * tok = mask[b,t] == -1 ? pad_idx : words[b, mask[b,t]] (words [Bq,nwords], mask [Bq,T] int64; lp_kind 0 =
 * plain fp32 rows, used to gather rows of the per-token input-projection table W_ih.emb + b) - replaces
 * get_srl_arg_seq_to_sent_seq + embed_tokens: code/mdl_vog.py:67-95, utils/mdl_srl_utils.py:124-127.
 * vog_lang_gather: out_lp[b*nsrl + s] = lp([full[cap[b,s,0]*Bq + b] | full[cap[b,s,1]*Bq + b]]), full
 * [T*Bq, D] fp32 time-major, cap [Bq,nsrl,2] int64 - the first/last-word gather + concat of
 * retrieve_srl_arg_from_lang_encode: code/mdl_vog.py:97-131.
 * vog_mask_rows: out[r] = x[r] * msk[r] (+ low-precision copy), msk int64 - the srl_arg_inds_msk product:
 * code/mdl_vog.py:137-140. */
int vog_lang_embed(const int64_t* words, int nwords, const int64_t* mask, int T, const float* emb, int E,
                   int64_t pad_idx, int Bq, void* out_lp, int lp_kind, void* stream);
int vog_lang_gather(const float* full, int D, const int64_t* cap, int T, int Bq, int nsrl, void* out_lp,
                    int lp_kind, void* stream);
int vog_mask_rows(const float* x, const int64_t* msk, int rows, int D, float* out, void* out_lp, int lp_kind,
                  void* stream);

/* Grounding loss, forward (LossB_SPAT / LossB_TEMP): logits [B,nsrl,P] f32,
 * props [B,P,pdim] (x1,y1,x2,y2,...), gt [B,K,5], frm_mask [B,P,K] u8 (1 = frames differ / padded),
 * pnt_mask [B,P] u8, srl_boxes / srl_lens [B,nsrl,nb] int64 (gt-box index and 0/1 validity per argument slot),
 * arg_boxes_mask [B,nsrl], cmp_msk [B,ncmp], target_cmp [B] int64.  target(b,s,p) = max_i(IoU(p, gt[srl_boxes
 * [b,s,i]]) * frm|pnt mask * [vid(p) == target_cmp[b]] * srl_lens[b,s,i]) > 0.5 (bit-exact: IEEE fp32 in the
 * reference's operation order); loss[0] = mean over {arg_boxes_mask[b,s] * cmp_msk[b,vid(p)] != 0} of
 * BCE-with-logits(logit, target) * P * loss_lambda (plain mean when no argument has boxes).  targets
 * [B,nsrl,P] u8 may be NULL; workspace: vog_loss_workspace_bytes().  replaces
 * code/mdl_conc_single.py:191-311,342-433 + utils/box_utils.py:54-118.
 * `spat` is the layout mode: 0 = temp rows [vid][frame][prop], 1 = spat rows [frame][vid][prop], 2 = sep
 * (LossB_SEP grounding term, code/mdl_conc_sep.py:236-365): B = (query, video) pairs, ncmp = 1, target_cmp[b] = 0
 * for the target video of its query and anything else otherwise, cmp_msk [B,1]; the mean selects by cmp_msk alone
 * and arg_boxes_mask only decides masked-vs-plain mean. */
int64_t vog_loss_workspace_bytes(int B, int nsrl, int P);
int vog_loss_fwd(const float* logits, const float* props, int pdim, const float* gt, const uint8_t* frm_mask,
                 const uint8_t* pnt_mask, const int64_t* srl_boxes, const int64_t* srl_lens,
                 const int64_t* arg_boxes_mask, const int64_t* cmp_msk, const int64_t* target_cmp, int B, int nsrl,
                 int nb, int P, int K, int ncmp, int nppf, int spat, float loss_lambda, uint8_t* targets,
                 void* workspace, float* loss, void* stream);

/* Backward of vog_loss_fwd with respect to the logits (first link of the backward chain):
 * grad_logits[i] = grad_out[0] * d loss / d logits[i] = grad_out * coef * w_i * (sigmoid(logits_i) - targets_i), with the
 * selection mask w and coef = P * loss_lambda / count taken from the workspace the forward call filled (pass the SAME
 * workspace and the `targets` it wrote).  grad_out: device pointer to one float.  What torch autograd derives for
 * code/mdl_conc_single.py:277-311 / code/mdl_conc_sep.py:323-365. */
int vog_loss_bwd(const float* logits, const uint8_t* targets, const void* workspace, const float* grad_out, float* grad_logits,
                 int B, int nsrl, int P, void* stream);

/* Fused Adam step on flat fp32 buffers of n elements (parameters, gradients, first / second moments), one launch:
 * torch.optim.Adam semantics (no weight decay, no amsgrad) with `step` the 1-based step count and the bias corrections
 * evaluated in double on the host; the gradient is multiplied by grad_scale first (1/world after a summed all-reduce).
 * replaces torch.optim.Adam(betas=(0.9, 0.99)).step() over ~110 tensors: code/main_dist.py:55,
 * utils/trn_utils.py:505,799-803. */
int vog_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int64_t step, double grad_scale, void* stream);

/* Verb loss of LossB_SEP (code/mdl_conc_sep.py:418-434): vidf [n] video-level logits (n = B*ncmp), verb_cmp [n] int64
 * 0/1 targets, verb_cross_cmp_msk [n,m] int64; loss[0] = mean over rows with any mask entry set of
 * BCE-with-logits(vidf, verb_cmp) * loss_lambda (NaN when no row is selected, like the reference's empty mean). */
int vog_verb_loss_fwd(const float* vidf, const int64_t* verb_cmp, const int64_t* verb_cross_cmp_msk, int n, int m,
                      float loss_lambda, float* loss, void* stream);

/* Contrastive-sample concatenation on the device: per-video tensors feat [B,ncmp,nfrm*nppf,D],
 * seg [B,ncmp,nfrm,Ds], props [B,ncmp,nfrm*nppf,pdim] -> the single "concatenated video" the SPAT / TEMP models read.
 * spat=1: rows reordered [vid][frame][prop] -> [frame][vid][prop] (seg: [vid][frame] -> [frame][vid]) and props columns
 * 0,2 += shift*vid (shift = 720); spat=0: row order kept, props column 4 += shift*vid (shift = 10) - feat_out / seg_out
 * must be NULL then (the inputs already are the concatenated tensors).  Any of the three outputs may be NULL to skip it.
 * Bit-identical to the host code it replaces: reshuffle_boxes / process_props and the seg transpose,
 * code/dat_loader_simple.py:1067-1103,1147-1153,1196-1207 (SPAT), :1231-1252,1290-1292 (TEMP). */
int vog_concat_videos(const float* feat, int D, const float* seg, int Ds, const float* props, int pdim, float* feat_out,
                      float* seg_out, float* props_out, int B, int ncmp, int nfrm, int nppf, int spat, float shift,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * training step: backward of the path.  The reference's backward is torch autograd
 * over its forward (utils/trn_utils.py:500-505: mdl(batch) -> loss_fn -> loss.backward() -> optimizer.step());
 * every entry point below is the analytic gradient of the forward operation it cites.  Gradient buffers marked
 * "accumulated" must be zero-initialised by the caller (they are added to atomically).
 * ------------------------------------------------------------------------------------------- */

/* C[M,N] = epi(sum_k A(m,k) B(k,n)),  A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]  (one unit stride each);
 * epi: (+bias[n]); relu; (+C when accumulate).  Exact fp32.  dX = dY.W (sbk = ldw, sbn = 1) and dW = dY^T.X
 * (sam = 1, sak = ldy) of every nn.Linear on the path: code/transformer_code.py:57-60,80-81,169-172,180,186;
 * code/mdl_vog.py:182-193,202-207,224-230.  Deep reductions (accumulate, no bias/relu) are split over K. */
int vog_sgemm_strided(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                      const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu, int accumulate,
                      void* stream);

/* out[n] += sum_m x[m,n]   (accumulated): bias gradients of the nn.Linear layers. */
int vog_colsum_acc(const float* x, int64_t ldx, float* out, int64_t M, int N, void* stream);

/* g = dy * [act > 0] -> out (fp32, nullable; may alias dy) and / or out_lp (VOG_LP_*), dbias[n] += sum_m g (nullable,
 * accumulated).  act = the forward's post-ReLU activation, fp32 (act_kind 0 / 2) or bf16 (act_kind 1).
 * Gradient of nn.ReLU after nn.Linear: code/transformer_code.py:80-81, code/mdl_vog.py:202-207,224-230. */
int vog_relu_bwd(const float* dy, int64_t ldy, const void* act, int64_t lda, int act_kind, float* out, int64_t ldo,
                 void* out_lp, int64_t ldlp, int lp_kind, float* dbias, int64_t M, int N, void* stream);

/* LayerNorm backward of the post-LN ResidualBlock (code/transformer_code.py:21-31): x = saved pre-normalisation sum
 * [M,d]; dx (fp32, nullable) / dx_lp (VOG_LP_*, nullable) = gradient w.r.t. x (= w.r.t. the residual AND the branch);
 * dgamma, dbeta, dxsum [d] accumulated (dxsum = column sums of dx = bias gradient of the branch's last nn.Linear;
 * nullable). */
int vog_layernorm_bwd(const float* dy, int64_t ldy, const float* x, int64_t ldx, const float* gamma, float* dx,
                      int64_t lddx, void* dx_lp, int64_t ldlp, int lp_kind, float* dgamma, float* dbeta, float* dxsum,
                      int64_t M, int d, float eps, void* stream);

/* Attention backward, exact fp32, flash-style recompute from the saved log-sum-exp (nothing N x N is stored).
 * q,k,v / dq,dk,dv: [Bt*N, ld] / [Bt*N, ldg] with heads as column chunks (off, dh host arrays), out / dout the
 * forward output and its gradient, lse [Bt,H,N] from vog_attn_fwd_f32, delta [Bt,H,N] scratch.  Rank-1 bias:
 * da [Bt*nbox, H] and dbpe [H] accumulated; dense bias: ddense [Bt,N,N,H] written (nullable).
 * Gradient of RelAttention / Attention (code/transformer_code.py:41-50,136-160) and of the bias construction
 * relu(Linear(5,H)(p_i - p_j)) (code/mdl_vog.py:477-488, utils/mdl_srl_utils.py:30-69). */
int vog_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ld, const float* out, int64_t ldo,
                     const float* dout, int64_t lddo, const float* lse, float* delta, float* dq, float* dk, float* dv,
                     int64_t ldg, int Bt, int N, int H, const int* off, const int* dh, float inv_scale, int bias_mode,
                     const float* a, int nbox, const float* bpe, const float* dense, float* da, float* dbpe,
                     float* ddense, float drop_p, uint64_t seed, void* stream);

/* dW[h,c] += sum_rows da[row,h] * normalised(props[row, c])  (accumulated): gradient of vog_pe_project w.r.t. the
 * pe_{obj,mul}_sub_enc weight (code/mdl_vog.py:446-451,459-463,580-585); the boxes carry no gradient (:497,506,624). */
int vog_pe_project_bwd(const float* props, int ldp, const float* da, float* dW, int rows, int H, float vid_w,
                       float vid_h, float fdiv, void* stream);

/* Gradient of the multimodal token matrix w.r.t. its factors (token (b,f,s,p) = [vis | lang], code/mdl_vog.py:316-344,
 * 693-699): dvis [B*nfrm*nppf2, dv] = sum over the nsrl slots (written); dlang [B*nsrl, dl] += sum over frames and
 * proposals (accumulated).  dtok [B*nfrm*nsrl*nppf2, dv+dl] fp32. */
int vog_xmul_bwd(const float* dtok, float* dvis, float* dlang, int B, int nfrm, int nsrl, int nppf2, int dv, int dl,
                 void* stream);

/* Segment half of the prop|seg rows (code/mdl_conc_single.py:50-66,156-174): dseg[slot, c] = [x[slot*nppf, pe+c] > 0]
 * * sum_p dx[slot*nppf + p, pe+c]  (gradient of the replicated relu(seg_encoder) rows; written). */
int vog_seg_rep_bwd(const float* dx, const float* x, int ld, int pe, int se, int nppf, float* dseg, int64_t nslots,
                    void* stream);

/* Scorer tail backward (lin2[2] on relu(lin2[0]) + inverse regroup, code/mdl_vog.py:224-230,675-677,724-737):
 * dlogits [B,nsrl,nfrm*nppf2]; h [M,K] post-ReLU hidden (fp32 or bf16 by h_kind); dh (fp32, nullable) / dh_lp
 * (VOG_LP_*, nullable) [M,K]; dw2 [K], db2 [1], db1 [K] accumulated. */
int vog_lin2_bwd(const float* dlogits, const void* h, int64_t ldh, int h_kind, const float* w2, float* dh, void* dh_lp,
                 int lp_kind, float* dw2, float* db2, float* db1, int64_t M, int K, int nfrm, int nsrl, int nppf2,
                 void* stream);

/* Language-side glue backward (code/mdl_vog.py:97-140; utils/mdl_srl_utils.py:96-128): scatter of the first / last
 * word gather (dfull [T*Bq, D] accumulated) and of the embedding lookup (demb [V+1, E] accumulated; the padding row and
 * steps beyond lens get nothing). */
int vog_lang_gather_bwd(const float* dcat, int D, const int64_t* cap, int T, int Bq, int nsrl, float* dfull,
                        void* stream);
int vog_lang_embed_bwd(const int64_t* words, int nwords, const int64_t* mask, int T, const float* dx, int E,
                       int64_t pad_idx, int Bq, const int64_t* lens, float* demb, void* stream);

/* LSTM backward building blocks (nn.LSTM, bidirectional, packed sequences: utils/mdl_srl_utils.py:100-152).
 * hout [T*Bq, 2H] fp32 hidden states of the layer (time-major), lens [Bq].
 *   vog_lstm_hprev     hprev [T*Bq, 2H]: the state each step started from (t-1 forward, t+1 reverse, 0 at the ends)
 *   vog_lstm_scan      G [T*Bq, 8H] gate pre-activations (gx + hprev.W_hh^T, recomputed by a GEMM) -> acts
 *                      [T*Bq, 2, 6, H]: i, f, g, o, tanh(c_t), c_{t-1}
 *   vog_lstm_bwd_steps dout [T*Bq, 2H] -> dG [T*Bq, 8H] (zeros beyond lens); dh_{prev} = dG_t . W_hh is taken from the
 *                      TRANSPOSED recurrent weight whh_t [2,H,4H].  H = 1024, Bq <= 4 on a 148-SM device: ONE persistent
 *                      launch with the weights resident on chip (csrc/lstm_bwd.cu: every CTA multiplies the gate rows
 *                      it owns into a partial sum over all columns and the CTAs exchange those through self-tagged
 *                      records; whh [2,4H,H], the untransposed weight, is optional (NULL allowed) and only makes that
 *                      kernel's one-time weight load coalesced); otherwise T dependent launches that stream whh_t from
 *                      L2.  workspace:
 *                      vog_lstm_bwd_workspace_bytes(Bq, H) bytes, 16-byte aligned; Bq <= 8 per call. */
int vog_lstm_hprev(const float* hout, const int64_t* lens, float* hprev, int T, int Bq, int H, void* stream);
int vog_lstm_scan(const float* G, const int64_t* lens, float* acts, int T, int Bq, int H, void* stream);
int64_t vog_lstm_bwd_workspace_bytes(int Bq, int H);
int vog_lstm_bwd_steps(const float* dout, const float* acts, const float* whh_t, const float* whh, const int64_t* lens,
                       float* dG, void* workspace, int64_t workspace_bytes, int T, int Bq, int H, void* stream);

/* Weight packing of one attention block for the tensor-core entry points: wq, wk, wv, wo
 * fp32 [d,d] (nn.Linear layout) -> wqkv [3*H*dhp, d] with the rows of every head zero-padded to dhp (row
 * (which*H + h)*dhp + r) and wo_p [d, H*dhp] with the matching zero-padded columns, as bf16 (VOG_LP_BF16) or
 * tf32-rounded fp32 (VOG_LP_TF32).  dh[H] = torch.chunk head widths (code/transformer_code.py:169-186). */
int vog_pack_weights(const float* wq, const float* wk, const float* wv, const float* wo, int d, int H, const int* dh,
                     int dhp, int lp_kind, void* wqkv, void* wo_p, void* stream);

/* One workspace query for every entry point that takes a caller-provided workspace: op = VOG_WS_* ; the
 * dimensions a, b, c, d, e mean (M, N, K, tf32, BN) for VOG_WS_TC_GEMM, (Bt, N, H) for VOG_WS_TC_ATTN and
 * VOG_WS_TC_ATTN_BWD, (Bq, H) for VOG_WS_LSTM and VOG_WS_LSTM_BWD, (B, nsrl, P) for VOG_WS_LOSS; unused ones are ignored. */
enum { VOG_WS_TC_GEMM = 0, VOG_WS_TC_ATTN = 1, VOG_WS_TC_ATTN_BWD = 2, VOG_WS_LSTM = 3, VOG_WS_LOSS = 4, VOG_WS_LSTM_BWD = 5 };
int64_t vog_workspace_bytes(int op, int a, int b, int c, int d, int e);

/* ---- training step on the tensor cores (compute mode 'bf16') ---------------------------------------------------
 * vog_tc_attn_fwd_train: vog_tc_attn_fwd that also keeps lse [Bt,H,N] (log2-domain log-sum-exp of every score row -
 * all the backward needs instead of the N x N probabilities) and applies dropout with probability drop_p to the
 * probabilities after the softmax (code/transformer_code.py:153); the mask is a counter-based function of
 * (seed, sequence, head, query, key) that vog_tc_attn_bwd regenerates.  Rank-1 bias or none. */
int vog_tc_attn_fwd_train(const void* q, const void* k, const void* v, int Bt, int N, int H, int dhp, const int* dh,
                          float inv_scale, int bias_mode, const float* a, int nbox, const float* bpe, void* out,
                          int64_t ldo, int out_kind, void* workspace, int64_t workspace_bytes, float* lse, float drop_p,
                          uint64_t seed, void* stream);

/* Attention backward on tcgen05 (recompute from lse): q,k,v [Bt,H,N,dhp] bf16 as written by vog_tc_gemm_qkv, o / dout
 * [Bt*N, ld >= H*dhp] bf16 (forward output and its gradient, heads in padded slots), lse from the forward.  Writes
 * dqkv [Bt*N, ldg >= 3*H*dhp] bf16: dQ | dK | dV in the row order of the packed Wq|Wk|Wv operand (column
 * (which*H + h)*dhp + c).  Rank-1 bias: da [Bt*nbox, H] and dbpe [H] (fp32) are ACCUMULATED.  workspace:
 * vog_tc_attn_bwd_workspace_bytes() bytes, 256-byte aligned (the bf16 probabilities and score gradients
 * [Bt*H, Npad, Npad] live there between the kernels of this call).  Gradient of RelAttention / Attention
 * (code/transformer_code.py:41-50,136-160) and of relu(Linear(5,H)(p_i - p_j)) (code/mdl_vog.py:477-488). */
int64_t vog_tc_attn_bwd_workspace_bytes(int Bt, int N, int H);
int vog_tc_attn_bwd(const void* q, const void* k, const void* v, const void* o, int64_t ldo, const void* dout,
                    int64_t lddo, const float* lse, int Bt, int N, int H, int dhp, const int* dh, float inv_scale,
                    int bias_mode, const float* a, int nbox, const float* bpe, void* dqkv, int64_t ldg, float* da,
                    float* dbpe, void* workspace, int64_t workspace_bytes, float drop_p, uint64_t seed, void* stream);

/* Element-wise dropout of the training forward and its backward: out (fp32, nullable; may alias x) / out_lp
 * (VOG_LP_*, nullable) = x * keep / (1-p) + residual (nullable), keep = counter-based function of (seed, stream_id,
 * row, column) - calling it on the output gradient with the same ids is the backward.  nn.Dropout of the two
 * ResidualBlock branches (code/transformer_code.py:26,31) and the LSTM input / inter-layer / output dropouts
 * (utils/mdl_srl_utils.py:104,128,150). */
int vog_dropout(const float* x, int64_t ldx, const float* residual, int64_t ldr, float* out, int64_t ldo, void* out_lp,
                int64_t ldlp, int lp_kind, int64_t M, int N, float p, uint64_t seed, int stream_id, void* stream);

/* Weight gradient on tcgen05: C[N1,N2] (fp32, ldc) += A[K,N1]^T . B[K,N2], A / B bf16 row-major (the contraction
 * index is the slow index of both: dY and X as the forward / backward kernels leave them; nothing is transposed in
 * HBM).  C must be zero-initialised or hold a running gradient.  Replaces autograd's mm(dY^T, X) of every nn.Linear
 * on the path (code/transformer_code.py:57-60,80-81,169-172,180,186; code/mdl_vog.py:202-207,224-230). */
int vog_tc_gemm_tn(const void* A, int64_t lda, const void* B, int64_t ldb, int K, int N1, int N2, float* C,
                   int64_t ldc, void* stream);

/* ---- debug hooks (not part of the data path) ----------------------------------------------------
 * vog_debug_gemm_trace: device buffer of 8 int64 that receives clock64 stamps of CTA 0 of every
 * following vog_tc_gemm launch (entry, setup done, first TMA issued, first stage landed, last MMA
 * committed, accumulator visible to the epilogue, epilogue done, exit); NULL switches it off.
 * vog_debug_attn_prof: device buffer of 16 int64 for the per-phase cycle counters of one softmax
 * warp / the MMA issuer of vog_tc_attn_fwd (library built with -DVOG_ATTN_PROFILE). */
void vog_debug_gemm_trace(void* buf);
void vog_debug_pdl(int on);                   /* A/B: 0 = plain stream-ordered launches instead of programmatic dependent launches */
void vog_debug_lstm_exchange(int mode);        /* h_t exchange protocol: 4 automatic (default: 3 for 3-4 sequences, else 2), 2 self-tagged per-CTA records, 3 records + two hidden units per warp, 0 tagged 64-bit words, 1 per-CTA release flags; bits 8-23: poll back-off in ns (protocol 2) */
void vog_debug_lstm_bwd_resident(int on);      /* 0 = per-step LSTM backward launches even where the persistent kernel applies */
void vog_debug_lstm_trace(void* buf);          /* 8 int64: matvec, reduce, cell+publish, poll, barrier cycles, steps */
void vog_debug_attn_prof(void* buf);
void vog_debug_attn_cluster(int c);            /* v2 attention cluster size: 1, 2, 4, or 0 = automatic */
void vog_debug_attn_impl(int impl);            /* 1 = Q/P through shared memory, 2 = Q/P in tensor memory (default) */

#ifdef __cplusplus
}
#endif
#endif /* VOG_B200_H */
