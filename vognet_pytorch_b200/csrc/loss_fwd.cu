// Grounding loss, forward: LossB_SPAT / LossB_TEMP of the reference
// (code/mdl_conc_single.py:180-433) with bbox_overlaps_batch (utils/box_utils.py:61-118) in two kernels
// instead of ~25 library launches over [B,P,100] and [B,nsrl,P,100] temporaries:
//
//   loss_targets_kernel  one warp per proposal: IoU against the gt boxes named by every SRL argument (+1 pixel
//                        convention, multiplied by pad_frm_mask | pad_pnt_mask, zero-area rules), restricted to the
//                        target video, target = max_i(iou_i * len_i) > 0.5; BCE-with-logits per (argument,
//                        proposal) and the (argument has boxes) x (video valid) mask.  The IoU arithmetic uses
//                        non-contracted IEEE fp32 operations in the reference's order, so the boolean targets are
//                        bit-exact.
//   loss_reduce_kernel   deterministic masked mean (fixed summation order, fp64 accumulation) * P * loss_lambda.
#include "common.cuh"
#include "kernels.h"

namespace vog {

__global__ void __launch_bounds__(256)
loss_targets_kernel(const float* __restrict__ logits, const float* __restrict__ props, int pdim,
                    const float* __restrict__ gt, const unsigned char* __restrict__ frm_mask,
                    const unsigned char* __restrict__ pnt_mask, const long long* __restrict__ srl_boxes,
                    const long long* __restrict__ srl_lens, const long long* __restrict__ arg_boxes_mask,
                    const long long* __restrict__ cmp_msk, const long long* __restrict__ target_cmp,
                    int B, int nsrl, int nb, int P, int K, int ncmp, int nppf, int spat,
                    float* __restrict__ el, unsigned char* __restrict__ msk, unsigned char* __restrict__ tgt)
{
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)B * P) return;
    const int b = (int)(wid / P), p = (int)(wid % P);
    const float* pr = props + ((size_t)b * P + p) * pdim;
    const float ax1 = pr[0], ay1 = pr[1], ax2 = pr[2], ay2 = pr[3];
    const float ax = __fadd_rn(__fsub_rn(ax2, ax1), 1.f), ay = __fadd_rn(__fsub_rn(ay2, ay1), 1.f);
    const float a_area = __fmul_rn(ax, ay);
    const bool a_zero = (ax == 1.f) && (ay == 1.f);
    const int vid = spat == 1 ? (p / nppf) % ncmp : p / (P / ncmp);
    const bool on_target = (long long)vid == target_cmp[b];
    const unsigned char pm = pnt_mask[(size_t)b * P + p];

    // every (argument, box slot) pair on its own lane, 32 pairs per pass
    const int npairs = nsrl * nb;
    for (int s0 = 0; s0 < nsrl; ++s0) {
        // bits of this argument's box slots that exceed the threshold (nb <= 32 per pass chunk)
        bool hit = false;
        for (int i0 = 0; i0 < nb; i0 += 32) {
            const int i = i0 + lane;
            bool h = false;
            if (i < nb) {
                const size_t si = ((size_t)b * nsrl + s0) * nb + i;
                const long long k = srl_boxes[si];
                const float len = (float)srl_lens[si];
                float ov = 0.f;
                if (k >= 0 && k < K) {
                    const float* g = gt + ((size_t)b * K + k) * 5;
                    const float gx = __fadd_rn(__fsub_rn(g[2], g[0]), 1.f), gy = __fadd_rn(__fsub_rn(g[3], g[1]), 1.f);
                    const float g_area = __fmul_rn(gx, gy);
                    float iw = __fadd_rn(__fsub_rn(fminf(ax2, g[2]), fmaxf(ax1, g[0])), 1.f);
                    float ih = __fadd_rn(__fsub_rn(fminf(ay2, g[3]), fmaxf(ay1, g[1])), 1.f);
                    iw = iw < 0.f ? 0.f : iw;
                    ih = ih < 0.f ? 0.f : ih;
                    const float inter = __fmul_rn(iw, ih);
                    const float ua = __fsub_rn(__fadd_rn(a_area, g_area), inter);
                    ov = __fdiv_rn(inter, ua);
                    const unsigned char fm = frm_mask[((size_t)b * P + p) * K + k] | pm;
                    ov = __fmul_rn(ov, (float)fm);
                    if (gx == 1.f && gy == 1.f) ov = 0.f;
                    if (a_zero) ov = -1.f;
                    ov = __fmul_rn(ov, on_target ? 1.f : 0.f);
                }
                h = __fmul_rn(ov, len) > 0.5f;
            }
            hit |= __any_sync(0xffffffffu, h);
        }
        if (lane == 0) {
            const size_t o = ((size_t)b * nsrl + s0) * P + p;
            const float x = logits[o];
            const float t = hit ? 1.f : 0.f;
            // binary_cross_entropy_with_logits: (1 - t) x + m + log(exp(-m) + exp(-x - m)),  m = max(-x, 0)
            const float m = fmaxf(-x, 0.f);
            el[o] = (1.f - t) * x + m + logf(expf(-m) + expf(-x - m));
            // mode 2 (SEP, code/mdl_conc_sep.py:341-355): the video mask alone selects; the argument mask only decides
            // masked-vs-plain mean in the reduction
            msk[o] = (unsigned char)((spat == 2 || arg_boxes_mask[(size_t)b * nsrl + s0] != 0) &&
                                     (cmp_msk[(size_t)b * ncmp + vid] != 0));
            if (tgt) tgt[o] = hit ? 1 : 0;
        }
    }
    (void)npairs;
}

__global__ void __launch_bounds__(1024)
loss_reduce_kernel(const float* __restrict__ el, const unsigned char* __restrict__ msk,
                   const long long* __restrict__ arg_boxes_mask, int n_args, long long n, int P, float lambda,
                   float* __restrict__ loss, int sep, float* __restrict__ stats)
{
    __shared__ double s_sum[1024];
    __shared__ double s_all[1024];
    __shared__ long long s_cnt[1024];
    __shared__ int s_any;
    const int tid = threadIdx.x;
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (int i = tid; i < n_args; i += 1024)
        if (arg_boxes_mask[i] > 0) s_any = 1;
    double sum = 0.0, all = 0.0;
    long long cnt = 0;
    for (long long i = tid; i < n; i += 1024) {            // fixed order: deterministic result
        const double v = (double)el[i];
        all += v;
        if (msk[i]) { sum += v; ++cnt; }
    }
    s_sum[tid] = sum; s_all[tid] = all; s_cnt[tid] = cnt;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if (tid < off) { s_sum[tid] += s_sum[tid + off]; s_all[tid] += s_all[tid + off]; s_cnt[tid] += s_cnt[tid + off]; }
        __syncthreads();
    }
    if (tid == 0) {
        // code/mdl_conc_single.py:304-311,408-414: masked mean if any argument has boxes, plain mean otherwise
        // SEP multiplies by the video mask BEFORE the branch (code/mdl_conc_sep.py:355-363): its plain mean is the
        // masked sum over all elements
        const double mean = s_any ? s_sum[0] / (double)s_cnt[0] : (sep ? s_sum[0] : s_all[0]) / (double)n;
        loss[0] = (float)(mean * (double)P) * lambda;
        // for the backward: d loss / d el[i] = coef * (masked ? msk[i] : 1)
        stats[0] = (float)((double)P / (s_any ? (double)s_cnt[0] : (double)n)) * lambda;
        stats[1] = (s_any || sep) ? 1.f : 0.f;
    }
}

// workspace: el [n] f32 | msk [n] u8 | (16-byte aligned) stats {coef, masked} f32
static long long loss_stats_offset(long long n) { return (n * 5 + 15) / 16 * 16; }
long long loss_workspace_bytes(int B, int nsrl, int P) { return loss_stats_offset((long long)B * nsrl * P) + 16; }

int loss_fwd(const float* logits, const float* props, int pdim, const float* gt, const unsigned char* frm_mask,
             const unsigned char* pnt_mask, const long long* srl_boxes, const long long* srl_lens,
             const long long* arg_boxes_mask, const long long* cmp_msk, const long long* target_cmp, int B,
             int nsrl, int nb, int P, int K, int ncmp, int nppf, int spat, float lambda, unsigned char* targets,
             void* workspace, float* loss, cudaStream_t st)
{
    VOG_REQUIRE(B > 0 && nsrl > 0 && nb > 0 && P > 0 && K > 0 && ncmp > 0 && P % ncmp == 0, "loss_fwd: bad dimension");
    VOG_REQUIRE(spat >= 0 && spat <= 2, "loss_fwd: mode must be 0 (temp), 1 (spat) or 2 (sep)");
    VOG_REQUIRE(spat != 1 || (nppf > 0 && P % (ncmp * nppf) == 0), "loss_fwd: spat grouping needs P %% (ncmp*nppf) == 0");
    VOG_REQUIRE(spat != 2 || ncmp == 1, "loss_fwd: sep mode takes one video per (query, video) pair");
    const long long n = (long long)B * nsrl * P;
    float* el = reinterpret_cast<float*>(workspace);
    unsigned char* msk = reinterpret_cast<unsigned char*>(workspace) + n * 4;
    const long long warps = (long long)B * P;
    loss_targets_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(logits, props, pdim, gt, frm_mask, pnt_mask, srl_boxes,
                                                                    srl_lens, arg_boxes_mask, cmp_msk, target_cmp, B, nsrl,
                                                                    nb, P, K, ncmp, nppf, spat, el, msk, targets);
    if (check_launch("loss_targets")) return -1;
    float* stats = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + loss_stats_offset(n));
    loss_reduce_kernel<<<1, 1024, 0, st>>>(el, msk, arg_boxes_mask, B * nsrl, n, P, lambda, loss, spat == 2, stats);
    return check_launch("loss_reduce");
}

// Backward of the grounding loss with respect to the logits (first link of SURVEY.md section 8f row 2): the loss is
// coef * sum_i w_i * BCE(x_i, t_i) with w_i the selection mask (or 1 for the plain mean), so
//     d loss / d x_i = grad_out * coef * w_i * (sigmoid(x_i) - t_i)
// - what autograd derives for binary_cross_entropy_with_logits + masked_select + mean in the reference
// (code/mdl_conc_single.py:277-311).  Reads the mask and the reduction statistics the forward left in its workspace.
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ tgt, const unsigned char* __restrict__ msk,
                const float* __restrict__ stats, const float* __restrict__ grad_out, float* __restrict__ grad, long long n)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float coef = stats[0] * grad_out[0];
    const float w = (stats[1] != 0.f) ? (float)msk[i] : 1.f;
    const float x = logits[i];
    const float sg = 1.f / (1.f + expf(-x));
    grad[i] = coef * w * (sg - (float)tgt[i]);
}

int loss_bwd(const float* logits, const unsigned char* targets, const void* workspace, const float* grad_out, float* grad,
             int B, int nsrl, int P, cudaStream_t st)
{
    const long long n = (long long)B * nsrl * P;
    if (n == 0) return 0;
    const unsigned char* ws = reinterpret_cast<const unsigned char*>(workspace);
    loss_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, targets, ws + n * 4,
                                                                reinterpret_cast<const float*>(ws + loss_stats_offset(n)),
                                                                grad_out, grad, n);
    return check_launch("loss_bwd");
}

// Verb loss of LossB_SEP (code/mdl_conc_sep.py:418-434): BCE-with-logits of the video-level logits against verb_cmp,
// mean over the (query, video) pairs whose verb_cross_cmp_msk row has any entry set (NaN when there is none, like the
// mean of an empty selection).  n = B*ncmp pairs, m = row length of the mask.
__global__ void verb_loss_kernel(const float* __restrict__ vidf, const long long* __restrict__ verb_cmp,
                                 const long long* __restrict__ vcc, int n, int m, float lambda, float* __restrict__ loss)
{
    __shared__ double s_sum[256];
    __shared__ int s_cnt[256];
    const int tid = threadIdx.x;
    double sum = 0.0; int cnt = 0;
    for (int i = tid; i < n; i += 256) {
        double rs = 0.0;                                    // float sum of 0/1 entries: exact
        for (int j = 0; j < m; ++j) rs += (double)(float)vcc[(size_t)i * m + j];
        if (rs > 0.0) {
            const float x = vidf[i], t = (float)verb_cmp[i];
            const float mx = fmaxf(-x, 0.f);
            sum += (double)((1.f - t) * x + mx + logf(expf(-mx) + expf(-x - mx)));
            ++cnt;
        }
    }
    s_sum[tid] = sum; s_cnt[tid] = cnt;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) { s_sum[tid] += s_sum[tid + off]; s_cnt[tid] += s_cnt[tid + off]; }
        __syncthreads();
    }
    if (tid == 0) loss[0] = (float)(s_sum[0] / (double)s_cnt[0]) * lambda;
}

int verb_loss_fwd(const float* vidf, const long long* verb_cmp, const long long* vcc, int n, int m, float lambda,
                  float* loss, cudaStream_t st)
{
    VOG_REQUIRE(n > 0 && m > 0, "verb_loss_fwd: empty input");
    verb_loss_kernel<<<1, 256, 0, st>>>(vidf, verb_cmp, vcc, n, m, lambda, loss);
    return check_launch("verb_loss");
}

}  // namespace vog
