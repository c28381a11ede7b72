// Backward kernels of the VOGNet fusion path, exact fp32 on CUDA cores, and the element-wise / reduction
// backward kernels shared by every compute mode (the tcgen05 backward uses the GEMM / attention kernels of
// tc_gemm_tn.cu / tc_attn_bwd.cu and these for everything else).
//
// The reference has no backward code of its own: its training step is torch autograd over the forward
// (utils/trn_utils.py:500-505: mdl(batch) -> loss_fn -> loss.backward() -> optimizer.step()).  Each kernel below
// is the analytic gradient of the forward operation cited at its entry point; parity is pinned on the gradients
// torch autograd derives for the UNMODIFIED reference (tests/golden/grad_cpu_ref.npz, oracle/make_golden.py).
#include "common.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace vog {

// =============================================================================================
// generic strided SGEMM:  C[m,n] = epi( sum_k A(m,k) B(k,n) ),  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
//   epi: (+ bias[n]) ; relu ; (+ C_old when accumulate)
// serves  dX = dY . W   (A = dY row-major, B = W [N_out, K_in] row-major: sbk = ldw, sbn = 1)
//         dW = dY^T . X (A(m,k) = dY[k, m]: sam = 1, sak = ldy;  B = X row-major)
//         and the forward NT form (B(k,n) = W[n, k]: sbk = 1, sbn = ldw).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles (same shape as sgemm_nt_kernel).
// =============================================================================================
constexpr int GS_BM = 64, GS_BN = 64, GS_BK = 16;

__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                     long long sbk, long long sbn, const float* __restrict__ bias, float* __restrict__ C,
                     long long ldc, int M, int N, int K, int relu, int accumulate, int kchunk)
{
    __shared__ float As[GS_BK][GS_BM + 4];
    __shared__ float Bs[GS_BK][GS_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // loader mappings chosen so that consecutive threads walk the contiguous dimension of each operand
    const bool a_kfast = sak == 1, b_nfast = sbn == 1;
    for (int k0 = kbeg; k0 < kend; k0 += GS_BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = tid + 256 * q;
            int m, k;
            if (a_kfast) { k = e % GS_BK; m = e / GS_BK; } else { m = e % GS_BM; k = e / GS_BM; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < kend) ? A[(long long)gm * sam + (long long)gk * sak] : 0.f;
            int n, kb;
            if (b_nfast) { n = e % GS_BN; kb = e / GS_BN; } else { kb = e % GS_BK; n = e / GS_BK; }
            const int gn = n0 + n, gkb = k0 + kb;
            Bs[kb][n] = (gn < N && gkb < kend) ? B[(long long)gkb * sbk + (long long)gn * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GS_BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            float* c = C + (long long)m * ldc + n;
            if (gridDim.z > 1) { atomicAdd(c, v); continue; }       // split-K: C pre-zeroed (or holds the addend)
            if (bias) v += bias[n];
            if (relu) v = fmaxf(v, 0.f);
            if (accumulate) v += *c;
            *c = v;
        }
    }
}

int sgemm_strided(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                  const float* bias, float* C, long long ldc, int M, int N, int K, int relu, int accumulate,
                  cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(sam == 1 || sak == 1, "sgemm_strided: A needs one unit stride");
    VOG_REQUIRE(sbk == 1 || sbn == 1, "sgemm_strided: B needs one unit stride");
    dim3 grid(cdiv(N, GS_BN), cdiv(M, GS_BM), 1);
    int kchunk = K > 0 ? K : 1;
    // deep reductions into small outputs (weight gradients: K = rows of the activation matrix): split K over
    // grid.z with fp32 atomics.  Only legal when the epilogue is a pure accumulation.
    const long long tiles = (long long)grid.x * grid.y;
    if (accumulate && !bias && !relu && K >= 2048 && tiles < 2LL * num_sms()) {
        int splits = (int)((4LL * num_sms()) / tiles);
        if (splits > K / 256) splits = K / 256;
        if (splits > 1) {
            kchunk = round_up(cdiv(K, splits), GS_BK);
            grid.z = cdiv(K, kchunk);
        }
    }
    sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, bias, C, ldc, M, N, K, relu, accumulate, kchunk);
    return check_launch("sgemm_strided");
}

// =============================================================================================
// out[n] += sum_m x[m, n] * (gate ? gate[m,n] > 0 : 1)        (bias gradients)      out is accumulated atomically
// =============================================================================================
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long ldx, float* __restrict__ out, long long M, int N, int rows_per_cta)
{
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    const long long r1 = min(M, r0 + rows_per_cta);
    float s = 0.f;
    if (c < N)
        for (long long r = r0 + w; r < r1; r += 8) s += x[r * ldx + c];
    red[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

int colsum_acc(const float* x, long long ldx, float* out, long long M, int N, cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    int rpc = 512;
    dim3 grid(cdiv(N, 32), (unsigned)((M + rpc - 1) / rpc));
    colsum_kernel<<<grid, 256, 0, st>>>(x, ldx, out, M, N, rpc);
    return check_launch("colsum");
}

// =============================================================================================
// ReLU backward (+ bias gradient):  g[m,n] = dy[m,n] * (act[m,n] > 0), written in place / as fp32 and / or as the
// low-precision A operand of the following tensor-core GEMMs;  dbias[n] += sum_m g[m,n] (atomic, optional).
// act is the forward's post-ReLU activation (fp32, or the bf16 / tf32 copy the forward kept).
// derivative of nn.ReLU after nn.Linear: code/transformer_code.py:80-81, code/mdl_vog.py:202-207,224-230.
// =============================================================================================
__device__ __forceinline__ float act_at(const void* act, int kind, long long idx) {
    return kind == 1 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(act)[idx])
                     : reinterpret_cast<const float*>(act)[idx];
}
__device__ __forceinline__ void store_lp1(void* base, long long idx, float v, int kind) {
    if (kind == 1) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else if (kind == 2) reinterpret_cast<float*>(base)[idx] = to_tf32(v);
    else reinterpret_cast<float*>(base)[idx] = v;
}

__global__ void __launch_bounds__(256)
relu_bwd_kernel(const float* __restrict__ dy, long long ldy, const void* __restrict__ act, long long lda, int act_kind,
                float* __restrict__ out, long long ldo, void* __restrict__ out_lp, long long ldlp, int lp_kind,
                float* __restrict__ dbias, long long M, int N, int rows_per_cta)
{
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    const long long r1 = min(M, r0 + rows_per_cta);
    float s = 0.f;
    if (c < N) {
        for (long long r = r0 + w; r < r1; r += 8) {
            const float a = act_at(act, act_kind, r * lda + c);
            const float g = a > 0.f ? dy[r * ldy + c] : 0.f;
            if (out) out[r * ldo + c] = g;
            if (out_lp) store_lp1(out_lp, r * ldlp + c, g, lp_kind);
            s += g;
        }
    }
    if (dbias) {
        red[w][threadIdx.x & 31] = s;
        __syncthreads();
        if (w == 0 && c < N) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
            atomicAdd(dbias + c, t);
        }
    }
}

int relu_bwd(const float* dy, long long ldy, const void* act, long long lda, int act_kind, float* out, long long ldo,
             void* out_lp, long long ldlp, int lp_kind, float* dbias, long long M, int N, cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(act_kind == 0 || act_kind == 1 || act_kind == 2, "relu_bwd: bad act_kind");
    const int rpc = 256;
    dim3 grid(cdiv(N, 32), (unsigned)((M + rpc - 1) / rpc));
    relu_bwd_kernel<<<grid, 256, 0, st>>>(dy, ldy, act, lda, act_kind, out, ldo, out_lp, ldlp, lp_kind, dbias, M, N, rpc);
    return check_launch("relu_bwd");
}

// =============================================================================================
// LayerNorm backward (nn.LayerNorm(d_model), eps inside the sqrt; post-LN ResidualBlock
// code/transformer_code.py:21-31):  y = (x - mean) * rstd * gamma + beta over the last dimension.
//   xhat = (x - mean) * rstd,  g = dy * gamma,
//   dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),   dgamma += sum_rows dy * xhat,   dbeta += sum_rows dy,
//   dxsum += sum_rows dx   (= bias gradient of the nn.Linear whose output the residual block normalises)
// x is the saved PRE-normalisation sum (residual + branch); mean / rstd are recomputed (two-pass, as the forward).
// One warp per row, the row lives in registers; every CTA walks a slab of rows, keeps its column sums in registers
// and publishes them once with atomics.
// =============================================================================================
template <int NC>     // columns per lane: d <= 32 * NC
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, long long ldy, const float* __restrict__ x, long long ldx,
                     const float* __restrict__ gamma, float* __restrict__ dx, long long lddx,
                     void* __restrict__ dx_lp, long long ldlp, int lp_kind, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, float* __restrict__ dxsum, long long M, int d, float eps,
                     int rows_per_cta)
{
    extern __shared__ float red[];                    // [8][d]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = min(M, r0 + rows_per_cta);
    float ag[NC], ab[NC], as_[NC], gm[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        ag[i] = ab[i] = as_[i] = 0.f;
        const int c = lane + 32 * i;
        gm[i] = c < d ? gamma[c] : 0.f;
    }
    const float inv_d = 1.f / d;
    for (long long r = r0 + w; r < r1; r += 8) {
        float xv[NC], gv[NC];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            xv[i] = c < d ? x[r * ldx + c] : 0.f;
            gv[i] = c < d ? dy[r * ldy + c] : 0.f;
            s += xv[i];
        }
        const float mean = warp_sum(s) * inv_d;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            xv[i] = c < d ? xv[i] - mean : 0.f;
            q = fmaf(xv[i], xv[i], q);
        }
        const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            xv[i] *= rstd;                             // xhat
            ag[i] = fmaf(gv[i], xv[i], ag[i]);
            ab[i] += gv[i];
            gv[i] *= gm[i];                            // g = dy * gamma
            m1 += gv[i];
            m2 = fmaf(gv[i], xv[i], m2);
        }
        m1 = warp_sum(m1) * inv_d;
        m2 = warp_sum(m2) * inv_d;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            const float v = rstd * (gv[i] - m1 - xv[i] * m2);
            if (c < d) {
                if (dx) dx[r * lddx + c] = v;
                if (dx_lp) store_lp1(dx_lp, r * ldlp + c, v, lp_kind);
                as_[i] += v;
            }
        }
    }
    // CTA-level column sums: three passes over one [8][d] buffer
    float* outs[3] = {dgamma, dbeta, dxsum};
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        if (outs[pass] == nullptr) continue;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const int c = lane + 32 * i;
            if (c < d) red[w * d + c] = pass == 0 ? ag[i] : (pass == 1 ? ab[i] : as_[i]);
        }
        __syncthreads();
        for (int c = threadIdx.x; c < d; c += 256) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i * d + c];
            atomicAdd(outs[pass] + c, t);
        }
    }
}

int layernorm_bwd(const float* dy, long long ldy, const float* x, long long ldx, const float* gamma, float* dx,
                  long long lddx, void* dx_lp, long long ldlp, int lp_kind, float* dgamma, float* dbeta, float* dxsum,
                  long long M, int d, float eps, cudaStream_t st)
{
    if (M == 0) return 0;
    VOG_REQUIRE(d >= 1 && d <= 1024, "layernorm_bwd: d=%d out of range (<= 1024)", d);
    // enough CTAs to fill the GPU a few times over, few enough that the atomics stay negligible
    long long rpc = (M + 4LL * 148 - 1) / (4LL * 148);
    if (rpc < 8) rpc = 8;
    rpc = (rpc + 7) / 8 * 8;
    const unsigned grid = (unsigned)((M + rpc - 1) / rpc);
    const size_t smem = (size_t)8 * d * sizeof(float);
#define VOG_LNB(NC)                                                                                                   \
    layernorm_bwd_kernel<NC><<<grid, 256, smem, st>>>(dy, ldy, x, ldx, gamma, dx, lddx, dx_lp, ldlp, lp_kind, dgamma,  \
                                                      dbeta, dxsum, M, d, eps, (int)rpc)
    if (d <= 256) VOG_LNB(8);
    else if (d <= 512) VOG_LNB(16);
    else if (d <= 768) VOG_LNB(24);
    else VOG_LNB(32);
#undef VOG_LNB
    return check_launch("layernorm_bwd");
}

// =============================================================================================
// delta[bt,h,i] = sum_c dO[i, off_h + c] * O[i, off_h + c]      (softmax backward row term)
// O / dO: [Bt*N, ld] fp32 with heads as column chunks (exact path) - one warp per (row, head)
// =============================================================================================
struct HeadSplit { int off[VOG_MAX_HEADS]; int dh[VOG_MAX_HEADS]; };

__global__ void __launch_bounds__(256)
attn_delta_f32_kernel(const float* __restrict__ o, long long ldo, const float* __restrict__ dout, long long lddo,
                      float* __restrict__ delta, int Bt, int N, int H, const HeadSplit hs)
{
    const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (wid >= (long long)Bt * N * H) return;
    const int h = (int)(wid % H);
    const long long row = wid / H;                    // bt*N + i
    const int off = hs.off[h], dh = hs.dh[h];
    float s = 0.f;
    for (int c = lane; c < dh; c += 32) s = fmaf(o[row * ldo + off + c], dout[row * lddo + off + c], s);
    s = warp_sum(s);
    if (lane == 0) {
        const long long bt = row / N, i = row % N;
        delta[(bt * H + h) * N + i] = s;
    }
}

// =============================================================================================
// attention backward, fp32, flash-style recompute (nothing N x N is stored):
//   z_ij = (q_i.k_j + b_ij) * c,  P = softmax_j z = exp(z - lse_i),  O = P V
//   dP_ij = dO_i . v_j,  dz_ij = P_ij (dP_ij - delta_i),  ds_ij = db_ij = c dz_ij
//   dq_i = sum_j ds_ij k_j,   dk_j = sum_i ds_ij q_i,   dv_j = sum_i P_ij dO_i
//   rank-1 bias b_ij = relu(a_i - a_j + bpe):  da_i += sum_j db_ij [b_ij > 0],  da_j -= sum_i db_ij [b_ij > 0],
//   dbpe += sum_ij db_ij [b_ij > 0]
// Two passes of ONE kernel (no atomics on dq / dk / dv, deterministic): the q-pass owns 32 query rows and streams
// key tiles, the kv-pass owns 32 key rows and streams query tiles.  Gradient of RelAttention / Attention
// (code/transformer_code.py:41-50,136-160) and of the bias construction (code/mdl_vog.py:477-488).
// =============================================================================================
struct AttnBwdF32Params {
    const float* q; const float* k; const float* v; long long ld;
    const float* dout; long long lddo;
    const float* lse; const float* delta;           // [Bt,H,N]
    float* dq; float* dk; float* dv; long long ldg;
    int Bt, N, H;
    int off[VOG_MAX_HEADS]; int dh[VOG_MAX_HEADS];
    float inv_scale;
    int bias_mode; const float* a; int nbox; const float* bpe; const float* dense;
    float* da; float* dbpe; float* ddense;
    int kv_pass;
    float drop_p; unsigned long long seed;       // the forward's dropout on the probabilities, regenerated
};

template <int KPT>
__global__ void __launch_bounds__(256)
attn_bwd_f32_kernel(const AttnBwdF32Params p)
{
    extern __shared__ float sm[];
    const int h = blockIdx.y, bt = blockIdx.z, r0 = blockIdx.x * 32;
    const int dh = p.dh[h], off = p.off[h], N = p.N;
    const int ldk = dh + 1;
    float* R1 = sm;                              // [32][dh]   q-pass: Q rows     kv-pass: K rows
    float* R2 = R1 + 32 * dh;                    // [32][dh]   q-pass: dO rows    kv-pass: V rows
    float* C1 = R2 + 32 * dh;                    // [64][dh+1] q-pass: K tile     kv-pass: Q tile
    float* C2 = C1 + 64 * ldk;                   // [64][dh+1] q-pass: V tile     kv-pass: dO tile
    float* Ss = C2 + 64 * ldk;                   // [32][65]   ds
    float* Ps = Ss + 32 * 65;                    // [32][65]   P (kv-pass)
    __shared__ float bsum_s[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long rowbase = (long long)bt * N;
    const bool kv = p.kv_pass != 0;
    const float* r1src = kv ? p.k : p.q;
    const float* r2src = kv ? p.v : p.dout;
    const long long r1ld = p.ld, r2ld = kv ? p.ld : p.lddo;
    const float* c1src = kv ? p.q : p.k;
    const float* c2src = kv ? p.dout : p.v;
    const long long c1ld = p.ld, c2ld = kv ? p.lddo : p.ld;
    for (int e = tid; e < 32 * dh; e += 256) {
        const int i = e / dh, c = e % dh;
        const int gi = r0 + i;
        R1[e] = gi < N ? r1src[(rowbase + gi) * r1ld + off + c] : 0.f;
        R2[e] = gi < N ? r2src[(rowbase + gi) * r2ld + off + c] : 0.f;
    }
    const long long statbase = ((long long)bt * p.H + h) * N;
    // per-row constants of the 4 rows this warp owns
    float r_lse[4], r_delta[4], r_a[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int gi = r0 + warp * 4 + r;
        r_lse[r] = r_delta[r] = r_a[r] = 0.f;
        if (gi < N) {
            if (!kv) { r_lse[r] = p.lse[statbase + gi]; r_delta[r] = p.delta[statbase + gi]; }
            if (p.bias_mode == 1) {
                r_a[r] = p.a[((long long)bt * p.nbox + gi % p.nbox) * p.H + h];
                if (!kv) r_a[r] += p.bpe[h];
            }
        }
    }
    float bias_acc[4] = {0.f, 0.f, 0.f, 0.f};     // sum over columns of db_ij [b_ij > 0] for this warp's rows
    const int orow = tid >> 3, ocol = tid & 7;
    float acc1[KPT], acc2[KPT];
#pragma unroll
    for (int k = 0; k < KPT; ++k) acc1[k] = acc2[k] = 0.f;

    for (int c0 = 0; c0 < N; c0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * dh; e += 256) {
            const int j = e / dh, c = e % dh;
            const int gj = c0 + j;
            C1[j * ldk + c] = gj < N ? c1src[(rowbase + gj) * c1ld + off + c] : 0.f;
            C2[j * ldk + c] = gj < N ? c2src[(rowbase + gj) * c2ld + off + c] : 0.f;
        }
        __syncthreads();
        float s[4][2], dp[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r][0] = s[r][1] = dp[r][0] = dp[r][1] = 0.f;
        {
            const float* a0 = C1 + lane * ldk; const float* a1 = C1 + (lane + 32) * ldk;
            const float* b0 = C2 + lane * ldk; const float* b1 = C2 + (lane + 32) * ldk;
            const float* r1p = R1 + warp * 4 * dh; const float* r2p = R2 + warp * 4 * dh;
            for (int c = 0; c < dh; ++c) {
                const float x0 = a0[c], x1 = a1[c], y0 = b0[c], y1 = b1[c];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float u = r1p[r * dh + c], w = r2p[r * dh + c];
                    s[r][0] = fmaf(u, x0, s[r][0]); s[r][1] = fmaf(u, x1, s[r][1]);
                    dp[r][0] = fmaf(w, y0, dp[r][0]); dp[r][1] = fmaf(w, y1, dp[r][1]);
                }
            }
        }
        // per-column constants
        float c_lse[2] = {0.f, 0.f}, c_delta[2] = {0.f, 0.f}, c_a[2] = {0.f, 0.f};
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int gj = c0 + lane + 32 * cc;
            if (gj < N) {
                if (kv) { c_lse[cc] = p.lse[statbase + gj]; c_delta[cc] = p.delta[statbase + gj]; }
                if (p.bias_mode == 1) {
                    c_a[cc] = p.a[((long long)bt * p.nbox + gj % p.nbox) * p.H + h];
                    if (kv) c_a[cc] += p.bpe[h];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int gi = r0 + warp * 4 + r;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int gj = c0 + lane + 32 * cc;
                const int qi = kv ? gj : gi, kj = kv ? gi : gj;          // query / key index of this element
                float pval = 0.f, ds = 0.f;
                if (gi < N && gj < N) {
                    float b = 0.f;
                    bool gate = false;
                    if (p.bias_mode == 1) {
                        const float t = kv ? c_a[cc] - r_a[r] : r_a[r] - c_a[cc];   // a_q + bpe - a_k
                        gate = t > 0.f;
                        b = fmaxf(t, 0.f);
                    } else if (p.bias_mode == 2) {
                        b = p.dense[(((long long)bt * N + qi) * N + kj) * p.H + h];
                    }
                    const float z = (s[r][cc] + b) * p.inv_scale;
                    const float lse = kv ? c_lse[cc] : r_lse[r];
                    const float dl = kv ? c_delta[cc] : r_delta[r];
                    pval = expf(z - lse);
                    float dpe = dp[r][cc];
                    if (p.drop_p > 0.f) {
                        uint32_t rnd[4];
                        attn_rand8x16(p.seed, (uint32_t)(bt * p.H + h), (uint32_t)qi, (uint32_t)(kj & ~15), rnd);
                        const bool keep = ((rnd[(kj & 15) >> 2] >> (8 * (kj & 3))) & 0xffu) >= drop_threshold8(p.drop_p);
                        const float ik = drop_inv_keep8(p.drop_p);
                        dpe = keep ? dpe * ik : 0.f;
                        ds = pval * (dpe - dl) * p.inv_scale;
                        pval = keep ? pval * ik : 0.f;           // what multiplied V in the forward (dV = P_drop^T dO)
                    } else
                    ds = pval * (dpe - dl) * p.inv_scale;
                    if (gate) bias_acc[r] += ds;
                    if (!kv && p.bias_mode == 2 && p.ddense)
                        p.ddense[(((long long)bt * N + qi) * N + kj) * p.H + h] = ds;
                }
                Ss[(warp * 4 + r) * 65 + lane + 32 * cc] = ds;
                if (kv) Ps[(warp * 4 + r) * 65 + lane + 32 * cc] = pval;
            }
        }
        __syncthreads();
        {
            const float* dsr = Ss + orow * 65;
            const float* pr = Ps + orow * 65;
            for (int j = 0; j < 64; ++j) {
                const float dj = dsr[j];
                const float* c1r = C1 + j * ldk + ocol;
#pragma unroll
                for (int k = 0; k < KPT; ++k)
                    if (ocol + 8 * k < dh) acc1[k] = fmaf(dj, c1r[8 * k], acc1[k]);
                if (kv) {
                    const float pj = pr[j];
                    const float* c2r = C2 + j * ldk + ocol;
#pragma unroll
                    for (int k = 0; k < KPT; ++k)
                        if (ocol + 8 * k < dh) acc2[k] = fmaf(pj, c2r[8 * k], acc2[k]);
                }
            }
        }
    }
    const int gi = r0 + orow;
    if (gi < N) {
        float* d1 = (kv ? p.dk : p.dq) + (rowbase + gi) * p.ldg + off;
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            const int c = ocol + 8 * k;
            if (c < dh) d1[c] = acc1[k];
        }
        if (kv) {
            float* d2 = p.dv + (rowbase + gi) * p.ldg + off;
#pragma unroll
            for (int k = 0; k < KPT; ++k) {
                const int c = ocol + 8 * k;
                if (c < dh) d2[c] = acc2[k];
            }
        }
    }
    if (p.bias_mode == 1) {
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float v = warp_sum(bias_acc[r]);
            const int g = r0 + warp * 4 + r;
            if (lane == 0 && g < N && v != 0.f)
                atomicAdd(p.da + ((long long)bt * p.nbox + g % p.nbox) * p.H + h, kv ? -v : v);
            tot += v;
        }
        if (!kv) {
            if (lane == 0) bsum_s[warp] = tot;
            __syncthreads();
            if (tid == 0) {
                float t = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) t += bsum_s[i];
                atomicAdd(p.dbpe + h, t);
            }
        }
    }
}

int attn_bwd_f32(const float* q, const float* k, const float* v, long long ld, const float* out, long long ldo,
                 const float* dout, long long lddo, const float* lse, float* delta, float* dq, float* dk, float* dv,
                 long long ldg, int Bt, int N, int H, const int* off, const int* dh, float inv_scale, int bias_mode,
                 const float* a, int nbox, const float* bpe, const float* dense, float* da, float* dbpe, float* ddense,
                 cudaStream_t st, float drop_p, unsigned long long seed)
{
    VOG_REQUIRE(H >= 1 && H <= VOG_MAX_HEADS, "attn_bwd_f32: H=%d out of range", H);
    VOG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attn_bwd_f32: dropout probability %f", (double)drop_p);
    VOG_REQUIRE(Bt <= 65535, "attn_bwd_f32: Bt=%d exceeds grid.z", Bt);
    if (Bt == 0 || N == 0) return 0;
    AttnBwdF32Params p;
    p.q = q; p.k = k; p.v = v; p.ld = ld; p.dout = dout; p.lddo = lddo; p.lse = lse; p.delta = delta;
    p.dq = dq; p.dk = dk; p.dv = dv; p.ldg = ldg; p.Bt = Bt; p.N = N; p.H = H; p.inv_scale = inv_scale;
    p.bias_mode = bias_mode; p.a = a; p.nbox = nbox > 0 ? nbox : 1; p.bpe = bpe; p.dense = dense;
    p.da = da; p.dbpe = dbpe; p.ddense = ddense;
    p.drop_p = drop_p; p.seed = seed;
    int dhmax = 0;
    HeadSplit hs = {};
    for (int h = 0; h < H; ++h) {
        p.off[h] = off[h]; p.dh[h] = dh[h];
        hs.off[h] = off[h]; hs.dh[h] = dh[h];
        dhmax = dh[h] > dhmax ? dh[h] : dhmax;
    }
    VOG_REQUIRE(dhmax <= 256, "attn_bwd_f32: head dim %d > 256 unsupported", dhmax);
    VOG_REQUIRE(bias_mode != 1 || (a && bpe && da && dbpe), "attn_bwd_f32: rank-1 bias needs a, bpe, da, dbpe");
    VOG_REQUIRE(bias_mode != 2 || dense, "attn_bwd_f32: dense bias pointer missing");
    {
        const long long nw = (long long)Bt * N * H;
        attn_delta_f32_kernel<<<(unsigned)((nw + 7) / 8), 256, 0, st>>>(out, ldo, dout, lddo, delta, Bt, N, H, hs);
        if (check_launch("attn_delta_f32")) return -1;
    }
    const size_t smem = sizeof(float) * (2 * 32 * dhmax + 2 * 64 * (dhmax + 1) + 2 * 32 * 65);
    dim3 grid(cdiv(N, 32), H, Bt);
    for (int pass = 0; pass < 2; ++pass) {
        p.kv_pass = pass;
#define VOG_LAUNCH_ATTB(KPT)                                                                                          \
    do {                                                                                                              \
        VOG_CUDA(cudaFuncSetAttribute(attn_bwd_f32_kernel<KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        attn_bwd_f32_kernel<KPT><<<grid, 256, smem, st>>>(p);                                                         \
    } while (0)
        if (dhmax <= 64) VOG_LAUNCH_ATTB(8);
        else if (dhmax <= 128) VOG_LAUNCH_ATTB(16);
        else if (dhmax <= 192) VOG_LAUNCH_ATTB(24);
        else VOG_LAUNCH_ATTB(32);
#undef VOG_LAUNCH_ATTB
        if (check_launch("attn_bwd_f32")) return -1;
    }
    return 0;
}

// =============================================================================================
// rank-1 bias factor backward: a[row,h] = W_h . norm(box_row)   (pe_project_kernel, code/mdl_vog.py:446-451,459-463)
//   dW[h,c] += sum_rows da[row,h] * norm(box_row)[c]
// =============================================================================================
__global__ void __launch_bounds__(256)
pe_project_bwd_kernel(const float* __restrict__ props, int ldp, const float* __restrict__ da, float* __restrict__ dW,
                      int rows, int H, float vw, float vh, float fdiv)
{
    __shared__ float red[8][VOG_MAX_HEADS * 5];
    float acc[VOG_MAX_HEADS * 5];
#pragma unroll
    for (int i = 0; i < VOG_MAX_HEADS * 5; ++i) acc[i] = 0.f;
    for (int row = blockIdx.x * 256 + threadIdx.x; row < rows; row += gridDim.x * 256) {
        const float* pr = props + (size_t)row * ldp;
        const float nb[5] = {pr[0] / vw, pr[1] / vh, pr[2] / vw, pr[3] / vh, pr[4] / fdiv};
#pragma unroll
        for (int h = 0; h < VOG_MAX_HEADS; ++h) {
            if (h < H) {
                const float g = da[(size_t)row * H + h];
#pragma unroll
                for (int c = 0; c < 5; ++c) acc[h * 5 + c] = fmaf(g, nb[c], acc[h * 5 + c]);
            }
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < VOG_MAX_HEADS * 5; ++i) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) red[w][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < H * 5) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(dW + threadIdx.x, t);
    }
}

int pe_project_bwd(const float* props, int ldp, const float* da, float* dW, int rows, int H, float vw, float vh,
                   float fdiv, cudaStream_t st)
{
    if (rows == 0) return 0;
    VOG_REQUIRE(H >= 1 && H <= VOG_MAX_HEADS, "pe_project_bwd: H=%d out of range", H);
    int grid = cdiv(rows, 256);
    if (grid > 296) grid = 296;
    pe_project_bwd_kernel<<<grid, 256, 0, st>>>(props, ldp, da, dW, rows, H, vw, vh, fdiv);
    return check_launch("pe_project_bwd");
}

// =============================================================================================
// multimodal token gradient -> its two factors.  Token (b, f, s, p) = [vis[(b*nfrm+f)*nppf2 + p] | lang[b*nsrl+s]]
// (concate_vis_lang_feats + conc_encode2 regroup, code/mdl_vog.py:316-344,693-699), so
//   dvis[(b*nfrm+f)*nppf2 + p, :dv] = sum_s dtok[(b,f,s,p), :dv]        (written, optionally += dvis_add)
//   dlang[b*nsrl + s, :]           += sum_{f,p} dtok[(b,f,s,p), dv:]    (atomic accumulation, pre-zeroed)
// =============================================================================================
__global__ void __launch_bounds__(256)
xmul_bwd_vis_kernel(const float* __restrict__ dtok, int dtot, float* __restrict__ dvis, int dv, int nsrl, int nppf2,
                    long long nvis)
{
    const int dv4 = dv / 4;
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= nvis * dv4) return;
    const long long row = idx / dv4;                  // bt*nppf2 + p
    const int c = (int)(idx % dv4) * 4;
    const long long bt = row / nppf2, pp = row % nppf2;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sl = 0; sl < nsrl; ++sl) {
        const float4 t = *reinterpret_cast<const float4*>(dtok + ((bt * nsrl + sl) * nppf2 + pp) * dtot + c);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(dvis + row * dv + c) = s;
}

__global__ void __launch_bounds__(256)
xmul_bwd_lang_kernel(const float* __restrict__ dtok, int dtot, float* __restrict__ dlang, int dv, int dl, int nfrm,
                     int nsrl, int nppf2, int chunk)
{
    // grid (B*nsrl, nfrm, ceil(nppf2/chunk)); thread = one language column
    const int bs = blockIdx.x, f = blockIdx.y;
    const int b = bs / nsrl, sl = bs % nsrl;
    const int p0 = blockIdx.z * chunk, p1 = min(nppf2, p0 + chunk);
    for (int c = threadIdx.x; c < dl; c += 256) {
        float s = 0.f;
        const float* base = dtok + ((((long long)b * nfrm + f) * nsrl + sl) * nppf2) * dtot + dv + c;
        for (int pp = p0; pp < p1; ++pp) s += base[(long long)pp * dtot];
        atomicAdd(dlang + (long long)bs * dl + c, s);
    }
}

int xmul_bwd(const float* dtok, float* dvis, float* dlang, int B, int nfrm, int nsrl, int nppf2, int dv, int dl,
             cudaStream_t st)
{
    if (B == 0) return 0;
    VOG_REQUIRE(dv % 4 == 0 && (dv + dl) % 4 == 0, "xmul_bwd: widths must be multiples of 4");
    const long long nvis = (long long)B * nfrm * nppf2;
    const long long n = nvis * (dv / 4);
    xmul_bwd_vis_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dtok, dv + dl, dvis, dv, nsrl, nppf2, nvis);
    if (check_launch("xmul_bwd_vis")) return -1;
    const int chunk = 64;
    dim3 grid(B * nsrl, nfrm, cdiv(nppf2, chunk));
    xmul_bwd_lang_kernel<<<grid, 256, 0, st>>>(dtok, dv + dl, dlang, dv, dl, nfrm, nsrl, nppf2, chunk);
    return check_launch("xmul_bwd_lang");
}

// =============================================================================================
// prop|seg row backward.  Row (b, vf, p) of the object-transformer input is [relu(prop) | relu(seg[b,vf])]
// (concat_prop_seg_feats, code/mdl_conc_single.py:50-66,156-174): the segment half was replicated over the nppf
// proposals of its (frame, video) slot, so its gradient is the sum over them (ReLU mask applied once, it is the same
// for every replica):   dseg[b*nvf + vf, c] = [x[row0, pe + c] > 0] * sum_p dx[(b,vf,p), pe + c]
// =============================================================================================
__global__ void __launch_bounds__(256)
seg_rep_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ x, int ld, int pe, int se, int nppf,
                   float* __restrict__ dseg, long long nslots)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= nslots * se) return;
    const long long slot = idx / se;
    const int c = (int)(idx % se);
    const long long row0 = slot * nppf;
    float s = 0.f;
    for (int pp = 0; pp < nppf; ++pp) s += dx[(row0 + pp) * ld + pe + c];
    dseg[slot * se + c] = x[row0 * ld + pe + c] > 0.f ? s : 0.f;
}

int seg_rep_bwd(const float* dx, const float* x, int ld, int pe, int se, int nppf, float* dseg, long long nslots,
                cudaStream_t st)
{
    if (nslots == 0) return 0;
    const long long n = nslots * se;
    seg_rep_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dx, x, ld, pe, se, nppf, dseg, nslots);
    return check_launch("seg_rep_bwd");
}

// =============================================================================================
// scorer tail backward.  logit[b,s,pidx] = w2 . h[m,:] + b2 with h = relu(lin2[0](token m)) and m = (b,f,s,p) the
// regrouped row of proposal pidx = f*nppf2 + p (code/mdl_vog.py:224-230,675-677,724-737):
//   dh[m,c] = dlogit * w2[c] * [h[m,c] > 0]     dw2[c] += dlogit * h[m,c]     db2 += dlogit
//   db1[c] += dh[m,c]                            (bias gradient of lin2[0])
// one warp per row m
// =============================================================================================
__global__ void __launch_bounds__(256)
lin2_bwd_kernel(const float* __restrict__ dlogits, const void* __restrict__ h, long long ldh, int h_kind,
                const float* __restrict__ w2, float* __restrict__ dh, void* __restrict__ dh_lp, int lp_kind,
                float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ db1, long long M, int K,
                int nfrm, int nsrl, int nppf2, int rows_per_cta)
{
    extern __shared__ float red[];                    // [2][8][K]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float aw[8], ab[8];                               // K <= 256: 8 columns per lane
#pragma unroll
    for (int i = 0; i < 8; ++i) aw[i] = ab[i] = 0.f;
    float sb2 = 0.f;
    const long long P = (long long)nfrm * nppf2;
    for (long long m = r0 + w; m < r1; m += 8) {
        long long t = m;
        const int pp = (int)(t % nppf2); t /= nppf2;
        const int s_ = (int)(t % nsrl); t /= nsrl;
        const int f = (int)(t % nfrm);
        const long long b = t / nfrm;
        const float g = dlogits[(b * nsrl + s_) * P + (long long)f * nppf2 + pp];
        sb2 += g;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < K) {
                const float hv = act_at(h, h_kind, m * ldh + c);
                const float d = hv > 0.f ? g * w2[c] : 0.f;
                if (dh) dh[m * K + c] = d;
                if (dh_lp) store_lp1(dh_lp, m * K + c, d, lp_kind);
                aw[i] = fmaf(g, hv, aw[i]);
                ab[i] += d;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        if (c < K) { red[w * K + c] = aw[i]; red[(8 + w) * K + c] = ab[i]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < K; c += 256) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { t1 += red[i * K + c]; t2 += red[(8 + i) * K + c]; }
        atomicAdd(dw2 + c, t1);
        atomicAdd(db1 + c, t2);
    }
    // every lane of a warp accumulated the same g's: one lane per warp publishes
    if (lane == 0 && sb2 != 0.f) atomicAdd(db2, sb2);
}

int lin2_bwd(const float* dlogits, const void* h, long long ldh, int h_kind, const float* w2, float* dh, void* dh_lp,
             int lp_kind, float* dw2, float* db2, float* db1, long long M, int K, int nfrm, int nsrl, int nppf2,
             cudaStream_t st)
{
    if (M == 0) return 0;
    VOG_REQUIRE(K >= 1 && K <= 256, "lin2_bwd: hidden width %d out of range (<= 256)", K);
    long long rpc = (M + 2LL * 148 - 1) / (2LL * 148);
    if (rpc < 8) rpc = 8;
    rpc = (rpc + 7) / 8 * 8;
    const unsigned grid = (unsigned)((M + rpc - 1) / rpc);
    lin2_bwd_kernel<<<grid, 256, (size_t)16 * K * sizeof(float), st>>>(dlogits, h, ldh, h_kind, w2, dh, dh_lp, lp_kind,
                                                                        dw2, db2, db1, M, K, nfrm, nsrl, nppf2, (int)rpc);
    return check_launch("lin2_bwd");
}

// =============================================================================================
// language-side glue backward
// =============================================================================================
// lang_gather backward: cat[b*nsrl+s] = [full[cap0*Bq+b] | full[cap1*Bq+b]]  ->  dfull (pre-zeroed) += scatter
__global__ void __launch_bounds__(128)
lang_gather_bwd_kernel(const float* __restrict__ dcat, int D, const long long* __restrict__ cap, int T, int Bq,
                       int nsrl, float* __restrict__ dfull)
{
    const int row = blockIdx.x;
    const int b = row / nsrl;
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) {
        const int which = c / D, col = c - which * D;
        long long t = cap[(size_t)row * 2 + which];
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        atomicAdd(dfull + ((size_t)t * Bq + b) * D + col, dcat[(size_t)row * 2 * D + c]);
    }
}

int lang_gather_bwd(const float* dcat, int D, const long long* cap, int T, int Bq, int nsrl, float* dfull, cudaStream_t st)
{
    if (Bq * nsrl == 0) return 0;
    lang_gather_bwd_kernel<<<Bq * nsrl, 128, 0, st>>>(dcat, D, cap, T, Bq, nsrl, dfull);
    return check_launch("lang_gather_bwd");
}

// embedding backward: x[t*Bq+b] = emb[tok(b,t)]  ->  demb[tok] += dx[t*Bq+b]  (padding row gets no gradient:
// nn.Embedding(padding_idx), utils/mdl_srl_utils.py:96-98)
__global__ void __launch_bounds__(128)
lang_embed_bwd_kernel(const long long* __restrict__ words, int nwords, const long long* __restrict__ mask, int T,
                      const float* __restrict__ dx, int E, long long pad_idx, int Bq, const long long* __restrict__ lens,
                      float* __restrict__ demb)
{
    const int row = blockIdx.x;                 // t*Bq + b
    const int t = row / Bq, b = row % Bq;
    if (lens && t >= lens[b]) return;           // beyond the packed length: never fed to the LSTM
    const long long mk = mask[(size_t)b * T + t];
    long long tok = pad_idx;
    if (mk != -1) tok = words[(size_t)b * nwords + (mk < 0 ? 0 : (mk >= nwords ? nwords - 1 : mk))];
    if (tok < 0 || tok >= pad_idx) return;
    for (int c = threadIdx.x; c < E; c += blockDim.x) atomicAdd(demb + (size_t)tok * E + c, dx[(size_t)row * E + c]);
}

int lang_embed_bwd(const long long* words, int nwords, const long long* mask, int T, const float* dx, int E,
                   long long pad_idx, int Bq, const long long* lens, float* demb, cudaStream_t st)
{
    if (T * Bq == 0) return 0;
    lang_embed_bwd_kernel<<<T * Bq, 128, 0, st>>>(words, nwords, mask, T, dx, E, pad_idx, Bq, lens, demb);
    return check_launch("lang_embed_bwd");
}

// =============================================================================================
// LSTM backward (nn.LSTM, 2 layers, bidirectional, packed sequences: utils/mdl_srl_utils.py:100-152).
// The forward kernel keeps only the hidden states; everything else is recomputed here:
//   1. lstm_hprev: h_{t-1} of every step in processing order (forward: t-1, reverse: t+1, zero at the ends)
//   2. gate pre-activations for ALL steps at once: G = gx + hprev . W_hh^T   (a GEMM - the states are known)
//   3. lstm_scan: cell recurrence per hidden unit -> activated gates, tanh(c_t), c_{t-1}
//   4. lstm_bwd_step x T: the sequential part, dh_{t-1} = dG_t . W_hh
//   5. weight / input gradients as GEMMs over all steps (host side, sgemm_strided or the tensor-core GEMMs)
// =============================================================================================
__global__ void __launch_bounds__(256)
lstm_hprev_kernel(const float* __restrict__ hout, const long long* __restrict__ lens, float* __restrict__ hprev, int T,
                  int Bq, int H)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)T * Bq * 2 * H) return;
    const int u = (int)(idx % H);
    const int d = (int)((idx / H) % 2);
    const int b = (int)((idx / (2 * H)) % Bq);
    const int t = (int)(idx / ((long long)2 * H * Bq));
    const int len = (int)lens[b];
    float v = 0.f;
    if (t < len) {
        const int tp = d == 0 ? t - 1 : t + 1;
        if (tp >= 0 && tp < len) v = hout[((long long)tp * Bq + b) * 2 * H + (long long)d * H + u];
    }
    hprev[idx] = v;
}

int lstm_hprev(const float* hout, const long long* lens, float* hprev, int T, int Bq, int H, cudaStream_t st)
{
    const long long n = (long long)T * Bq * 2 * H;
    if (n == 0) return 0;
    lstm_hprev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hout, lens, hprev, T, Bq, H);
    return check_launch("lstm_hprev");
}

// acts [T*Bq, 2, 6, H]: i, f, g, o (activated), tanh(c_t), c_{t-1}; zeros at steps beyond the sequence length
__global__ void __launch_bounds__(256)
lstm_scan_kernel(const float* __restrict__ G, const long long* __restrict__ lens, float* __restrict__ acts, int T, int Bq,
                 int H)
{
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)Bq * 2 * H) return;
    const int u = (int)(idx % H);
    const int d = (int)((idx / H) % 2);
    const int b = (int)(idx / (2 * H));
    const int len = (int)lens[b];
    float c = 0.f;
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? s : T - 1 - s;
        const long long row = (long long)t * Bq + b;
        float* a = acts + ((row * 2 + d) * 6) * H + u;
        if (t >= len) {
#pragma unroll
            for (int k = 0; k < 6; ++k) a[(long long)k * H] = 0.f;
            continue;
        }
        const float* g = G + row * 8 * H + (long long)d * 4 * H + u;
        const float gi = 1.f / (1.f + expf(-g[0]));
        const float gf = 1.f / (1.f + expf(-g[H]));
        const float gg = tanhf(g[2 * H]);
        const float go = 1.f / (1.f + expf(-g[3 * H]));
        const float cp = c;
        c = fmaf(gf, c, gi * gg);
        a[0] = gi; a[H] = gf; a[2 * (long long)H] = gg; a[3 * (long long)H] = go;
        a[4 * (long long)H] = tanhf(c); a[5 * (long long)H] = cp;
    }
}

int lstm_scan(const float* G, const long long* lens, float* acts, int T, int Bq, int H, cudaStream_t st)
{
    const long long n = (long long)Bq * 2 * H;
    if (n == 0 || T == 0) return 0;
    lstm_scan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(G, lens, acts, T, Bq, H);
    return check_launch("lstm_scan");
}

// One backward timestep for both directions.  grid (H / LB_CH, 2), one warp per output hidden unit; every CTA
// recomputes the element-wise gate gradients of ALL hidden units of its direction (cheap, and the matvec needs all
// of them), writes the ones of its own chunk to dG and the carried cell gradient, then produces its chunk of
//     dh_{prev}[b, u'] = sum_r dG_t[b, r] * W_hh[r, u'],  r over the 4H gate rows
// from the TRANSPOSED recurrent weight whh_t [2][H][4H]: row u' is one contiguous 16 KB stream per warp (float4 per
// lane, fully coalesced, L2 resident after the first step) against the gate gradients in shared memory.
// carry buffers are double buffered by step parity: [2][2 dir][Bq][H].
constexpr int LB_CH = 16;          // hidden units (rows of W_hh^T) per CTA = warps per CTA: 64 CTAs x 2 directions at H = 1024
constexpr int LB_MAXB = 8;
constexpr int LB_THREADS = 32 * LB_CH;

__global__ void __launch_bounds__(LB_THREADS)
lstm_bwd_step_kernel(const float* __restrict__ dout, const float* __restrict__ acts, const float* __restrict__ whh_t,
                     const long long* __restrict__ lens, float* __restrict__ dG, float* __restrict__ dh_carry,
                     float* __restrict__ dc_carry, int step, int T, int Bq, int H)
{
    extern __shared__ __align__(16) float dgs[];     // [Bq][4H] gate-pre-activation gradients of this step
    const int d = blockIdx.y, u0 = blockIdx.x * LB_CH;
    const int t = d == 0 ? T - 1 - step : step;
    const int par = step & 1;
    const float* dh_in = dh_carry + ((long long)(par * 2 + d) * Bq) * H;
    const float* dc_in = dc_carry + ((long long)(par * 2 + d) * Bq) * H;
    float* dh_out = dh_carry + ((long long)((par ^ 1) * 2 + d) * Bq) * H;
    float* dc_out = dc_carry + ((long long)((par ^ 1) * 2 + d) * Bq) * H;
    for (int e = threadIdx.x; e < Bq * H; e += LB_THREADS) {
        const int b = e / H, u = e % H;
        const bool active = t < (int)lens[b];
        const long long row = (long long)t * Bq + b;
        float gi_ = 0.f, gf_ = 0.f, gg_ = 0.f, go_ = 0.f, dcn = dc_in[e];
        if (active) {
            const float* a = acts + ((row * 2 + d) * 6) * H + u;
            const float gi = a[0], gf = a[H], gg = a[2 * (long long)H], go = a[3 * (long long)H];
            const float tc = a[4 * (long long)H], cp = a[5 * (long long)H];
            const float dh = dout[row * 2 * H + (long long)d * H + u] + dh_in[e];
            const float dc = fmaf(dh * go, 1.f - tc * tc, dc_in[e]);
            go_ = dh * tc * go * (1.f - go);
            gi_ = dc * gg * gi * (1.f - gi);
            gg_ = dc * gi * (1.f - gg * gg);
            gf_ = dc * cp * gf * (1.f - gf);
            dcn = dc * gf;
        }
        float* s = dgs + (long long)b * 4 * H + u;
        s[0] = gi_; s[H] = gf_; s[2 * H] = gg_; s[3 * H] = go_;
        if (u >= u0 && u < u0 + LB_CH) {
            float* g = dG + row * 8 * H + (long long)d * 4 * H + u;
            g[0] = gi_; g[H] = gf_; g[2 * (long long)H] = gg_; g[3 * (long long)H] = go_;
            dc_out[e] = dcn;
        }
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int uo = u0 + w;
    const float4* wrow = reinterpret_cast<const float4*>(whh_t + ((long long)d * H + uo) * 4 * H);
    float acc[LB_MAXB];
#pragma unroll
    for (int b = 0; b < LB_MAXB; ++b) acc[b] = 0.f;
    for (int r4 = lane; r4 < H; r4 += 32) {          // 4H / 4 float4 per row
        const float4 wv = __ldg(wrow + r4);
#pragma unroll
        for (int b = 0; b < LB_MAXB; ++b)
            if (b < Bq) {
                const float4 g = *reinterpret_cast<const float4*>(dgs + (long long)b * 4 * H + 4 * r4);
                acc[b] = fmaf(g.x, wv.x, fmaf(g.y, wv.y, fmaf(g.z, wv.z, fmaf(g.w, wv.w, acc[b]))));
            }
    }
#pragma unroll
    for (int b = 0; b < LB_MAXB; ++b)
        if (b < Bq) {
            const float s = warp_sum(acc[b]);
            if (lane == 0) {
                const bool active = t < (int)lens[b];
                const long long e = (long long)b * H + uo;
                dh_out[e] = active ? s : dh_in[e];
            }
        }
}

int lstm_bwd_steps(const float* dout, const float* acts, const float* whh_t, const float* whh, const long long* lens, float* dG,
                   float* carry_ws, long long ws_bytes, int T, int Bq, int H, cudaStream_t st)
{
    if (T == 0 || Bq == 0) return 0;
    VOG_REQUIRE(ws_bytes >= (long long)8 * Bq * H * 4, "lstm_bwd_steps: workspace of %lld bytes is too small", ws_bytes);
    {   // H = 1024, <= 4 sequences, 148 SMs: one persistent weight-resident launch (lstm_bwd.cu)
        const int r = lstm_bwd_resident(dout, acts, whh_t, whh, lens, dG, carry_ws, ws_bytes, T, Bq, H, st);
        if (r != 0) return r < 0 ? -1 : 0;
    }
    VOG_REQUIRE(Bq <= LB_MAXB, "lstm_bwd_steps: at most %d sequences per call (got %d)", LB_MAXB, Bq);
    VOG_REQUIRE(H % LB_CH == 0 && H % 32 == 0, "lstm_bwd_steps: H=%d must be a multiple of 32", H);
    VOG_REQUIRE((reinterpret_cast<uintptr_t>(whh_t) & 15) == 0, "lstm_bwd_steps: whh_t must be 16-byte aligned");
    const size_t carry = (size_t)2 * 2 * Bq * H;
    VOG_CUDA(cudaMemsetAsync(carry_ws, 0, 2 * carry * sizeof(float), st));
    float* dh_carry = carry_ws;
    float* dc_carry = carry_ws + carry;
    const size_t smem = (size_t)Bq * 4 * H * sizeof(float);
    VOG_CUDA(cudaFuncSetAttribute(lstm_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(H / LB_CH, 2);
    for (int s = 0; s < T; ++s) {
        lstm_bwd_step_kernel<<<grid, LB_THREADS, smem, st>>>(dout, acts, whh_t, lens, dG, dh_carry, dc_carry, s, T, Bq, H);
        if (check_launch("lstm_bwd_step")) return -1;
    }
    return 0;
}


// =============================================================================================
// Element-wise dropout (nn.Dropout / F.dropout on a [M,N] activation): the two residual branches of a
// ResidualBlock (code/transformer_code.py:26,31) and the LSTM input / inter-layer / output dropouts
// (utils/mdl_srl_utils.py:104,128,150).  out = x * keep / (1-p) (+ residual); keep is a counter-based function of
// (seed, stream, row, column) - the backward calls the same kernel on the output gradient with the same ids.
// =============================================================================================
template <bool kVec>
__global__ void __launch_bounds__(256)
dropout_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ res, long long ldr,
               float* __restrict__ out, long long ldo, void* __restrict__ out_lp, long long ldlp, int lp_kind,
               long long M, int N, float p, unsigned long long seed, unsigned int stream)
{
    const int n8 = (N + 7) >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * n8) return;
    const long long r = idx / n8;
    const int c0 = (int)(idx - r * n8) * 8;
    uint32_t rnd[4];
    attn_rand16x8(seed, stream, (uint32_t)r, (uint32_t)c0, rnd);
    const uint32_t thr = drop_threshold16(p);
    const float inv_keep = 1.f / (1.f - p);
    if constexpr (kVec) {
        // N % 8 == 0 and 16-byte aligned rows (host-checked): two float4 per operand, one 32-byte segment per thread
        float v[8];
        {
            const float4 a = *reinterpret_cast<const float4*>(x + r * ldx + c0), b = *reinterpret_cast<const float4*>(x + r * ldx + c0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
            v[e] = ((rnd[e >> 1] >> (16 * (e & 1))) & 0xffffu) >= thr ? v[e] * inv_keep : 0.f;
        if (res) {
            const float4 a = *reinterpret_cast<const float4*>(res + r * ldr + c0), b = *reinterpret_cast<const float4*>(res + r * ldr + c0 + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        if (out) {
            *reinterpret_cast<float4*>(out + r * ldo + c0) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(out + r * ldo + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (out_lp) {
            if (lp_kind == 1) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out_lp) + r * ldlp + c0) =
                    make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                               *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) store_lp1(out_lp, r * ldlp + c0 + e, v[e], lp_kind);
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = c0 + e;
            if (c >= N) break;
            const bool keep = ((rnd[e >> 1] >> (16 * (e & 1))) & 0xffffu) >= thr;
            float v = keep ? x[r * ldx + c] * inv_keep : 0.f;
            if (res) v += res[r * ldr + c];
            if (out) out[r * ldo + c] = v;
            if (out_lp) store_lp1(out_lp, r * ldlp + c, v, lp_kind);
        }
    }
}

int dropout_apply(const float* x, long long ldx, const float* res, long long ldr, float* out, long long ldo, void* out_lp,
                  long long ldlp, int lp_kind, long long M, int N, float p, unsigned long long seed, unsigned int stream,
                  cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    VOG_REQUIRE(p >= 0.f && p < 1.f, "dropout: probability %f", (double)p);
    VOG_REQUIRE(M < (1LL << 32), "dropout: too many rows");
    const long long n = M * ((N + 7) >> 3);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = N % 8 == 0 && al16(x) && ldx % 4 == 0 && (!res || (al16(res) && ldr % 4 == 0)) &&
                     (!out || (al16(out) && ldo % 4 == 0)) && (!out_lp || (al16(out_lp) && ldlp % 8 == 0));
    if (vec)
        dropout_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, ldx, res, ldr, out, ldo, out_lp, ldlp, lp_kind, M,
                                                                          N, p, seed, stream);
    else
        dropout_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, ldx, res, ldr, out, ldo, out_lp, ldlp, lp_kind, M,
                                                                           N, p, seed, stream);
    return check_launch("dropout");
}

}  // namespace vog
