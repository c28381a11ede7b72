// Exact-fp32 CUDA-core kernels of the VOGNet fusion path (compute='fp32x': every product and sum
// in IEEE fp32, no tensor cores).  They are the numerically tight GPU path (parity vs the oracle at
// ~1e-5), the generic fallback for shapes/modes the tcgen05 kernels do not take (dense [Bt,N,N,H]
// bias tensors handed to RelTransformer.forward, odd head sizes) and the bring-up reference for
// the tensor-core kernels.  What each kernel replaces in the reference is cited at its entry
// point in vog_abi.cu / include/vog_b200.h.
#include "common.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace vog {

// =============================================================================================
// C[M,N] = epi(A[M,K] . W[N,K]^T)   64x64x16 tiles, 256 threads, 4x4 register micro-tiles
// epi: v = acc (+bias[n]) ; relu ; (+residual[m,n])
// =============================================================================================
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                const float* __restrict__ bias, const float* __restrict__ R, int ldr,
                float* __restrict__ C, int ldc, int M, int N, int K, int relu)
{
    __shared__ float As[SG_BK][SG_BM + 4];
    __shared__ float Ws[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    const int tx = tid % 16, ty = tid / 16;          // 16x16 threads, each 4x4 outputs
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread (one row-quad of k)
    const int lr = tid / 4, lk = (tid % 4) * 4;
    for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k = k0 + lk + q;
            int gm = m0 + lr, gn = n0 + lr;
            As[lk + q][lr] = (gm < M && k < K) ? A[(size_t)gm * lda + k] : 0.f;
            Ws[lk + q][lr] = (gn < N && k < K) ? W[(size_t)gn * ldw + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SG_BK; ++k) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = Ws[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (relu) v = fmaxf(v, 0.f);
            if (R) v += R[(size_t)m * ldr + n];
            C[(size_t)m * ldc + n] = v;
        }
    }
}

int sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias,
             const float* R, int ldr, float* C, int ldc, int M, int N, int K, int relu,
             cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    dim3 grid(cdiv(N, SG_BN), cdiv(M, SG_BM));
    sgemm_nt_kernel<<<grid, 256, 0, st>>>(A, lda, W, ldw, bias, R, ldr, C, ldc, M, N, K, relu);
    return check_launch("sgemm_nt");
}

// =============================================================================================
// multi-head attention with relative-position bias, fp32, flash-style (never materialises N x N)
//   out[bt,i,off_h+c] = sum_j softmax_j((q_i.k_j + bias_h(i,j)) * inv_scale) v_j[c]
// grid (ceil(N/32), H, Bt), 256 threads, 32 query rows x 64-key tiles.
// =============================================================================================
constexpr int AT_BQ = 32, AT_BKV = 64;

struct AttnF32Params {
    const float* q; const float* k; const float* v; int ld;      // [Bt*N, ld], head h at column off[h]
    float* out; int ldo;
    int Bt, N, H;
    int off[VOG_MAX_HEADS]; int dh[VOG_MAX_HEADS];
    float inv_scale;
    int bias_mode;              // 0 none, 1 rank-1 (a_i - a_j + b_h)+, 2 dense [Bt,N,N,H]
    const float* a; int nbox;   // mode 1: a [Bt*nbox, H]
    const float* bpe;           // mode 1: device [H]
    const float* dense;         // mode 2
    float* lse;                 // optional [Bt,H,N]: log-sum-exp of the scaled scores (saved for the backward)
    float drop_p;               // training: dropout on the probabilities (same counter-based mask as the tcgen05 kernels)
    unsigned long long seed;
};

template <int KPT>   // accumulator columns per thread: dh <= 8*KPT
__global__ void __launch_bounds__(256)
attn_f32_kernel(const AttnF32Params p)
{
    extern __shared__ float sm[];
    const int h = blockIdx.y, bt = blockIdx.z, q0 = blockIdx.x * AT_BQ;
    const int dh = p.dh[h], off = p.off[h], N = p.N;
    const int ldk = dh + 1;
    float* Qs = sm;                              // [32][dh]
    float* KVs = Qs + AT_BQ * dh;                // [64][dh+1]
    float* Ss = KVs + AT_BKV * ldk;              // [32][65]
    float* alpha_s = Ss + AT_BQ * 65;            // [32]
    float* linv_s = alpha_s + AT_BQ;             // [32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t rowbase = (size_t)bt * N;

    for (int e = tid; e < AT_BQ * dh; e += 256) {
        int i = e / dh, c = e % dh;
        int gi = q0 + i;
        Qs[e] = gi < N ? p.q[(rowbase + gi) * p.ld + off + c] : 0.f;
    }
    // softmax state of the 4 rows this warp owns (replicated over lanes)
    float m_run[4], l_run[4], ai[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        m_run[r] = -INFINITY; l_run[r] = 0.f;
        int gi = q0 + warp * 4 + r;
        ai[r] = (p.bias_mode == 1 && gi < N)
                    ? p.a[((size_t)bt * p.nbox + gi % p.nbox) * p.H + h] + p.bpe[h] : 0.f;
    }
    // PV ownership: row orow, columns ocol + 8*k
    const int orow = tid >> 3, ocol = tid & 7;
    float acc[KPT];
#pragma unroll
    for (int k = 0; k < KPT; ++k) acc[k] = 0.f;

    for (int j0 = 0; j0 < N; j0 += AT_BKV) {
        __syncthreads();                         // previous PV done with KVs / Ss
        for (int e = tid; e < AT_BKV * dh; e += 256) {
            int j = e / dh, c = e % dh;
            int gj = j0 + j;
            KVs[j * ldk + c] = gj < N ? p.k[(rowbase + gj) * p.ld + off + c] : 0.f;
        }
        __syncthreads();
        // ---- S = q.k^T for rows warp*4..+4, cols lane, lane+32
        float s[4][2];
#pragma unroll
        for (int r = 0; r < 4; ++r) s[r][0] = s[r][1] = 0.f;
        const float* k0p = KVs + lane * ldk;
        const float* k1p = KVs + (lane + 32) * ldk;
        const float* qp = Qs + warp * 4 * dh;
        for (int c = 0; c < dh; ++c) {
            float kk0 = k0p[c], kk1 = k1p[c];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float qq = qp[r * dh + c];
                s[r][0] = fmaf(qq, kk0, s[r][0]);
                s[r][1] = fmaf(qq, kk1, s[r][1]);
            }
        }
        // ---- bias, scale, online softmax
        float aj[2] = {0.f, 0.f};
        if (p.bias_mode == 1) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                int gj = j0 + lane + 32 * c;
                if (gj < N) aj[c] = p.a[((size_t)bt * p.nbox + gj % p.nbox) * p.H + h];
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int gi = q0 + warp * 4 + r;
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                int gj = j0 + lane + 32 * c;
                float v = s[r][c];
                if (p.bias_mode == 1) v += fmaxf(ai[r] - aj[c], 0.f);
                else if (p.bias_mode == 2 && gi < N && gj < N)
                    v += p.dense[(((size_t)bt * N + gi) * N + gj) * p.H + h];
                v *= p.inv_scale;
                if (gj >= N) v = -INFINITY;
                s[r][c] = v;
                mx = fmaxf(mx, v);
            }
            mx = warp_max(mx);
            float m_new = fmaxf(m_run[r], mx);
            float al = expf(m_run[r] - m_new);          // exp(-inf)=0 on the first tile
            float p0 = expf(s[r][0] - m_new), p1 = expf(s[r][1] - m_new);
            float rs = warp_sum(p0 + p1);
            l_run[r] = l_run[r] * al + rs;
            m_run[r] = m_new;
            if (p.drop_p > 0.f) {
                // dropout after the softmax (code/transformer_code.py:153): the row sum keeps every probability
                const uint32_t thr = drop_threshold8(p.drop_p);
                const float ik = drop_inv_keep8(p.drop_p);
                uint32_t rnd[4];
                const int g0 = j0 + lane, g1 = j0 + lane + 32;
                attn_rand8x16(p.seed, (uint32_t)(bt * p.H + h), (uint32_t)gi, (uint32_t)(g0 & ~15), rnd);
                p0 = ((rnd[(g0 & 15) >> 2] >> (8 * (g0 & 3))) & 0xffu) >= thr ? p0 * ik : 0.f;
                attn_rand8x16(p.seed, (uint32_t)(bt * p.H + h), (uint32_t)gi, (uint32_t)(g1 & ~15), rnd);
                p1 = ((rnd[(g1 & 15) >> 2] >> (8 * (g1 & 3))) & 0xffu) >= thr ? p1 * ik : 0.f;
            }
            Ss[(warp * 4 + r) * 65 + lane] = p0;
            Ss[(warp * 4 + r) * 65 + lane + 32] = p1;
            if (lane == 0) alpha_s[warp * 4 + r] = al;
        }
        __syncthreads();                         // S complete, K no longer needed
        for (int e = tid; e < AT_BKV * dh; e += 256) {
            int j = e / dh, c = e % dh;
            int gj = j0 + j;
            KVs[j * ldk + c] = gj < N ? p.v[(rowbase + gj) * p.ld + off + c] : 0.f;
        }
        __syncthreads();
        // ---- O = O*alpha + P.V
        float al = alpha_s[orow];
#pragma unroll
        for (int k = 0; k < KPT; ++k) acc[k] *= al;
        const float* pr = Ss + orow * 65;
        for (int j = 0; j < AT_BKV; ++j) {
            float pj = pr[j];
            const float* vr = KVs + j * ldk + ocol;
#pragma unroll
            for (int k = 0; k < KPT; ++k)
                if (ocol + 8 * k < dh) acc[k] = fmaf(pj, vr[8 * k], acc[k]);
        }
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            linv_s[warp * 4 + r] = 1.f / l_run[r];
            const int gr = q0 + warp * 4 + r;
            if (p.lse && gr < N) p.lse[((size_t)bt * p.H + h) * N + gr] = m_run[r] + logf(l_run[r]);
        }
    }
    __syncthreads();
    int gi = q0 + orow;
    if (gi < N) {
        float li = linv_s[orow];
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            int c = ocol + 8 * k;
            if (c < dh) p.out[(rowbase + gi) * p.ldo + off + c] = acc[k] * li;
        }
    }
}

int attn_f32(const float* q, const float* k, const float* v, int ld, float* out, int ldo,
             int Bt, int N, int H, const int* off, const int* dh, float inv_scale,
             int bias_mode, const float* a, int nbox, const float* bpe, const float* dense,
             cudaStream_t st, float* lse, float drop_p, unsigned long long seed)
{
    VOG_REQUIRE(H >= 1 && H <= VOG_MAX_HEADS, "attn_f32: H=%d out of range (max %d)", H, VOG_MAX_HEADS);
    VOG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attn_f32: dropout probability %f", (double)drop_p);
    VOG_REQUIRE(Bt <= 65535, "attn_f32: Bt=%d exceeds grid.z", Bt);
    if (Bt == 0 || N == 0) return 0;
    AttnF32Params p;
    p.q = q; p.k = k; p.v = v; p.ld = ld; p.out = out; p.ldo = ldo;
    p.Bt = Bt; p.N = N; p.H = H; p.inv_scale = inv_scale;
    p.bias_mode = bias_mode; p.a = a; p.nbox = nbox > 0 ? nbox : 1; p.dense = dense; p.bpe = bpe; p.lse = lse;
    p.drop_p = drop_p; p.seed = seed;
    int dhmax = 0;
    for (int h = 0; h < H; ++h) {
        p.off[h] = off[h]; p.dh[h] = dh[h];
        dhmax = dh[h] > dhmax ? dh[h] : dhmax;
    }
    VOG_REQUIRE(dhmax <= 256, "attn_f32: head dim %d > 256 unsupported", dhmax);
    VOG_REQUIRE(bias_mode != 1 || (a != nullptr && bpe != nullptr), "attn_f32: bias_mode 1 needs a and bpe");
    VOG_REQUIRE(bias_mode != 2 || dense != nullptr, "attn_f32: bias_mode 2 needs dense bias");
    size_t smem = sizeof(float) * (AT_BQ * dhmax + AT_BKV * (dhmax + 1) + AT_BQ * 65 + 2 * AT_BQ);
    dim3 grid(cdiv(N, AT_BQ), H, Bt);
#define VOG_LAUNCH_ATT(KPT)                                                                      \
    do {                                                                                         \
        VOG_CUDA(cudaFuncSetAttribute(attn_f32_kernel<KPT>,                                      \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        attn_f32_kernel<KPT><<<grid, 256, smem, st>>>(p);                                        \
    } while (0)
    if (dhmax <= 64) VOG_LAUNCH_ATT(8);
    else if (dhmax <= 128) VOG_LAUNCH_ATT(16);
    else if (dhmax <= 192) VOG_LAUNCH_ATT(24);
    else VOG_LAUNCH_ATT(32);
#undef VOG_LAUNCH_ATT
    return check_launch("attn_f32");
}

// =============================================================================================
// out = LayerNorm(x (+ r)) * w + b     one warp per row, two-pass statistics, eps inside sqrt
// optional low-precision copy for the next GEMM's A operand (bf16, or tf32-rounded fp32)
// =============================================================================================
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ r, int ldr,
                     const float* __restrict__ w, const float* __restrict__ b,
                     float* __restrict__ out, int ldo, void* __restrict__ out_lp, int ldlp, int lp_kind,
                     int M, int d, float eps)
{
    int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= M) return;
    const float* xr = x + (size_t)row * ldx;
    const float* rr = r ? r + (size_t)row * ldr : nullptr;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s += xr[c] + (rr ? rr[c] : 0.f);
    float mean = warp_sum(s) / d;
    float v = 0.f;
    for (int c = lane; c < d; c += 32) {
        float t = xr[c] + (rr ? rr[c] : 0.f) - mean;
        v = fmaf(t, t, v);
    }
    float rstd = rsqrtf(warp_sum(v) / d + eps);
    for (int c = lane; c < d; c += 32) {
        float t = (xr[c] + (rr ? rr[c] : 0.f) - mean) * rstd * w[c] + b[c];
        if (out) out[(size_t)row * ldo + c] = t;
        if (out_lp) {
            if (lp_kind == 1) ((__nv_bfloat16*)out_lp)[(size_t)row * ldlp + c] = __float2bfloat16_rn(t);
            else ((float*)out_lp)[(size_t)row * ldlp + c] = to_tf32(t);
        }
    }
}

// register-cached variant: d % 128 == 0, d <= 1024, 16-byte aligned rows - the row is read from
// global memory exactly once (float4 per lane, fully coalesced) and kept in registers for the
// mean / variance / normalise passes (two-pass variance as in the generic kernel)
template <int NV>
__global__ void __launch_bounds__(256)
add_layernorm_vec_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ r, int ldr,
                         const float* __restrict__ w, const float* __restrict__ b,
                         float* __restrict__ out, int ldo, void* __restrict__ out_lp, int ldlp, int lp_kind,
                         int M, float eps)
{
    constexpr int d = NV * 128;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();
    if (row >= M) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * ldx);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
    if (r) {
        const float4* rr = reinterpret_cast<const float4*>(r + (size_t)row * ldr);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 t = rr[lane + 32 * i];
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q);
        q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + lane + 32 * i);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
        float4 t;
        t.x = v[i].x * rstd * w4.x + b4.x; t.y = v[i].y * rstd * w4.y + b4.y;
        t.z = v[i].z * rstd * w4.z + b4.z; t.w = v[i].w * rstd * w4.w + b4.w;
        const size_t c = (size_t)(lane + 32 * i) * 4;
        if (out) *reinterpret_cast<float4*>(out + (size_t)row * ldo + c) = t;
        if (out_lp) {
            if (lp_kind == 1) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_lp) + (size_t)row * ldlp + c) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_lp) + (size_t)row * ldlp + c) =
                    make_float4(to_tf32(t.x), to_tf32(t.y), to_tf32(t.z), to_tf32(t.w));
            }
        }
    }
}

int add_layernorm(const float* x, int ldx, const float* r, int ldr, const float* w, const float* b,
                  float* out, int ldo, void* out_lp, int ldlp, int lp_kind, int M, int d, float eps,
                  cudaStream_t st)
{
    if (M == 0) return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = d % 128 == 0 && d <= 1024 && al16(x) && ldx % 4 == 0 && (!r || (al16(r) && ldr % 4 == 0)) &&
                     al16(w) && al16(b) && (!out || (al16(out) && ldo % 4 == 0)) &&
                     (!out_lp || ((reinterpret_cast<uintptr_t>(out_lp) & (lp_kind == 1 ? 7 : 15)) == 0 && ldlp % 4 == 0));
    if (vec) {
#define VOG_LN_CASE(NV)                                                                                   \
    case NV:                                                                                              \
        VOG_CUDA(launch_pdl(add_layernorm_vec_kernel<NV>, dim3(cdiv(M, 8)), dim3(256), 0, st, x, ldx, r, ldr, w, b, out, ldo, \
                            out_lp, ldlp, lp_kind, M, eps));                                              \
        break;
        switch (d / 128) {
            VOG_LN_CASE(1) VOG_LN_CASE(2) VOG_LN_CASE(3) VOG_LN_CASE(4) VOG_LN_CASE(5) VOG_LN_CASE(6)
            VOG_LN_CASE(7) VOG_LN_CASE(8)
        }
#undef VOG_LN_CASE
        return check_launch("add_layernorm_vec");
    }
    VOG_CUDA(launch_pdl(add_layernorm_kernel, dim3(cdiv(M, 8)), dim3(256), 0, st, x, ldx, r, ldr, w, b, out, ldo, out_lp, ldlp,
                        lp_kind, M, d, eps));
    return check_launch("add_layernorm");
}

// =============================================================================================
// a[row,h] = scale * W_h . (x1/vw, y1/vh, x2/vw, y2/vh, frame/fdiv)      rank-1 factor of the bias
// =============================================================================================
__global__ void pe_project_kernel(const float* __restrict__ props, int ldp, const float* __restrict__ W,
                                  float* __restrict__ a, int rows, int H, float vw, float vh, float fdiv,
                                  float scale)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= rows * H) return;
    int row = idx / H, h = idx % H;
    const float* p = props + (size_t)row * ldp;
    const float* w = W + h * 5;
    // same evaluation order as a 5-term dot product; the divisions are the reference's in-place
    // normalisation (code/mdl_vog.py:459-463)
    float acc = (p[0] / vw) * w[0];
    acc = fmaf(p[1] / vh, w[1], acc);
    acc = fmaf(p[2] / vw, w[2], acc);
    acc = fmaf(p[3] / vh, w[3], acc);
    acc = fmaf(p[4] / fdiv, w[4], acc);
    a[idx] = acc * scale;
}

// The same projection plus, in the same launch, the per-key factors the fused attention kernel consumes:
//   ak[bt*H + h, key] = c * a[(bt*nbox + key % nbox), h] for key < N, 0 up to the row end ld
// (what tc_attn's bias_expand pre-kernel would compute from `a`; every thread RE-derives its a value with the code above,
// so the two outputs agree bit for bit with pe_project + bias_expand).  rows == Bt * nbox.
__global__ void pe_project_expand_kernel(const float* __restrict__ props, int ldp, const float* __restrict__ W,
                                         float* __restrict__ a, float* __restrict__ ak, int rows, int H, float vw, float vh,
                                         float fdiv, float scale, int Bt, int N, int nbox, int ld, float c)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    auto project = [&](int row, int h) {
        const float* p = props + (size_t)row * ldp;
        const float* w = W + h * 5;
        float acc = (p[0] / vw) * w[0];
        acc = fmaf(p[1] / vh, w[1], acc);
        acc = fmaf(p[2] / vw, w[2], acc);
        acc = fmaf(p[3] / vh, w[3], acc);
        acc = fmaf(p[4] / fdiv, w[4], acc);
        return acc * scale;
    };
    if (idx < (long long)rows * H) a[idx] = project((int)(idx / H), (int)(idx % H));
    if (idx < (long long)Bt * H * ld) {
        const int key = (int)(idx % ld);
        const int bh = (int)(idx / ld);
        const int bt = bh / H, h = bh % H;
        ak[idx] = key < N ? project(bt * nbox + key % nbox, h) * c : 0.f;
    }
}

int pe_project_expand(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw, float vh,
                      float fdiv, float scale, float* ak, int Bt, int N, int nbox, int ld, float c, cudaStream_t st)
{
    if (rows == 0) return 0;
    VOG_REQUIRE(nbox > 0 && rows == Bt * nbox, "pe_project_expand: rows=%d != Bt*nbox=%d*%d", rows, Bt, nbox);
    const long long n = (long long)rows * H > (long long)Bt * H * ld ? (long long)rows * H : (long long)Bt * H * ld;
    VOG_CUDA(launch_pdl(pe_project_expand_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, props, ldp, W, a, ak,
                        rows, H, vw, vh, fdiv, scale, Bt, N, nbox, ld, c));
    return check_launch("pe_project_expand");
}

int pe_project(const float* props, int ldp, const float* W, float* a, int rows, int H, float vw,
               float vh, float fdiv, float scale, cudaStream_t st)
{
    if (rows == 0) return 0;
    int n = rows * H;
    VOG_CUDA(launch_pdl(pe_project_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, props, ldp, W, a, rows, H, vw, vh, fdiv, scale));
    return check_launch("pe_project");
}

// =============================================================================================
// selection (K4): scores [B,nsrl,P] -> per (b,s,frame,vid) max/argmax over nppf, gather box rows,
// argmax over vids.  Bit-exact vs torch.max / argmax: first (lowest) index wins ties, NaN wins.
// one warp per (b,s,frame,vid) group; one extra pass for the per-frame vid argmax.
// =============================================================================================
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
    // torch.max semantics: NaN propagates (first NaN wins), otherwise larger value, lowest index
    bool vn = v != v, bn = bv != bv;
    if (vn || bn) return vn && (!bn || i < bi);
    return v > bv || (v == bv && i < bi);
}

// one warp per (b, s, frame): the per-video proposal max / argmax + box gather for each of the ncmp videos, then the
// argmax over the videos (SPAT) - a single launch, no intermediate round trip
__global__ void __launch_bounds__(256)
select_kernel(const float* __restrict__ scores, const float* __restrict__ props, int pdim,
              float* __restrict__ boxes, float* __restrict__ out_scores, long long* __restrict__ indexs,
              int B, int nsrl, int ncmp, int nfrm, int nppf, int spat, const float* __restrict__ fin)
{
    // fin != nullptr: SEP layout - scores [B,ncmp,nsrl,nfrm*nppf], one proposal block per video, and the predicted
    // video is the argmax of the fused per-video score fin [B,ncmp] (code/eval_vsrl_corr.py:162-220)
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= B * nsrl * nfrm) return;
    const int frm = w % nfrm, bs = w / nfrm;            // bs = b*nsrl + s
    const int b = bs / nsrl;
    const int P = ncmp * nfrm * nppf;
    float vbest = 0.f; int vidx = 0;
    for (int vid = 0; vid < ncmp; ++vid) {
        const int base = spat ? (frm * ncmp + vid) * nppf : (vid * nfrm + frm) * nppf;
        const float* sc = fin ? scores + (((size_t)b * ncmp + vid) * nsrl + (bs - b * nsrl)) * (nfrm * nppf) + frm * nppf
                              : scores + (size_t)bs * P + base;
        float bv = 0.f; int bi = 0x7fffffff;
        bool have = false;
        for (int i = lane; i < nppf; i += 32) {
            float v = sc[i];
            if (!have || better(v, i, bv, bi)) { bv = v; bi = i; have = true; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            bool oh = __shfl_xor_sync(0xffffffffu, (int)have, o);
            if (oh && (!have || better(ov, oi, bv, bi))) { bv = ov; bi = oi; have = true; }
        }
        const size_t g = ((size_t)bs * ncmp + vid) * nfrm + frm;         // output order [b][s][vid][frm]
        if (lane == 0) out_scores[g] = bv;
        const float* pr = props + ((size_t)b * P + base + bi) * pdim;
        for (int c = lane; c < pdim; c += 32) boxes[g * pdim + c] = pr[c];
        if (vid == 0 || better(bv, vid, vbest, vidx)) { vbest = bv; vidx = vid; }
    }
    if (fin) {
        vbest = fin[(size_t)b * ncmp]; vidx = 0;
        for (int vid = 1; vid < ncmp; ++vid) {
            const float v = fin[(size_t)b * ncmp + vid];
            if (better(v, vid, vbest, vidx)) { vbest = v; vidx = vid; }
        }
    }
    if (lane == 0) indexs[w] = (spat || fin) ? vidx : 0;
}

int select_fwd(const float* scores, const float* props, int pdim, float* boxes, float* out_scores,
               long long* indexs, int B, int nsrl, int ncmp, int nfrm, int nppf, int spat,
               cudaStream_t st, const float* fin)
{
    const int nwarps = B * nsrl * nfrm;
    if (nwarps == 0) return 0;
    select_kernel<<<cdiv(nwarps, 8), 256, 0, st>>>(scores, props, pdim, boxes, out_scores, indexs, B, nsrl,
                                                   ncmp, nfrm, nppf, spat, fin);
    return check_launch("select");
}

// =============================================================================================
// SEP: fused per-video score (code/mdl_conc_sep.py:62-117, use_vis_msk).  One block per (query, video): one warp per
// SRL argument takes the max proposal logit (sigmoid is monotone: max of sigmoids = sigmoid of the max), the verb
// slot is replaced by the video-level verb score, then masked mean over the populated slots.
// =============================================================================================
__global__ void sep_fin_kernel(const float* __restrict__ logits, const float* __restrict__ vidf,
                               const long long* __restrict__ srl_msk, const long long* __restrict__ verb_ind,
                               const long long* __restrict__ cmp_msk, float* __restrict__ fin_loss,
                               float* __restrict__ fin_eval, int nsrl, int P1)
{
    __shared__ float best[32];
    const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nsrl) {
        const float* lg = logits + ((size_t)q * nsrl + warp) * P1;
        float m = -INFINITY; bool nan = false;
        for (int i = lane; i < P1; i += 32) { const float v = lg[i]; nan |= (v != v); m = fmaxf(m, v); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            nan |= (bool)__shfl_xor_sync(0xffffffffu, (int)nan, o);
        }
        if (lane == 0) best[warp] = nan ? NAN : 1.f / (1.f + expf(-m));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long vi = verb_ind[q];
        if (vi >= 0 && vi < nsrl) best[vi] = 1.f / (1.f + expf(-vidf[q]));
        const float cm = (float)cmp_msk[q];
        float sum = 0.f, cnt = 0.f;
        for (int s = 0; s < nsrl; ++s) {
            const float mk = (float)srl_msk[(size_t)q * nsrl + s];
            const float v = best[s] * mk;
            sum += v; cnt += mk;
            fin_loss[(size_t)q * nsrl + s] = v * cm;
        }
        fin_eval[q] = sum / cnt * cm;
    }
}

int sep_fin_scores(const float* logits, const float* vidf, const long long* srl_msk, const long long* verb_ind,
                   const long long* cmp_msk, float* fin_loss, float* fin_eval, int Bq, int nsrl, int P1, cudaStream_t st)
{
    if (Bq == 0) return 0;
    VOG_REQUIRE(nsrl >= 1 && nsrl <= 32 && P1 >= 1, "sep_fin_scores: nsrl=%d must be 1..32", nsrl);
    sep_fin_kernel<<<Bq, 32 * nsrl, 0, st>>>(logits, vidf, srl_msk, verb_ind, cmp_msk, fin_loss, fin_eval, nsrl, P1);
    return check_launch("sep_fin_scores");
}

// =============================================================================================
// fp32 -> bf16 / tf32-rounded copy of an activation matrix (A operand of the tcgen05 GEMMs)
// =============================================================================================
__global__ void cast_lp_kernel(const float* __restrict__ src, long long lds, void* __restrict__ dst,
                               long long ldd, long long rows, int cols4, int kind)
{
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cols4) return;
    long long r = idx / cols4; int c = (int)(idx % cols4) * 4;
    float4 v = *reinterpret_cast<const float4*>(src + r * lds + c);
    if (kind == 1) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst) + r * ldd + c) = o;
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + r * ldd + c) =
            make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
}

int cast_lp(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols, int kind,
            cudaStream_t st)
{
    if (rows == 0 || cols == 0) return 0;
    VOG_REQUIRE(cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0, "cast_lp: cols and leading dims must be multiples of 4");
    VOG_REQUIRE(kind == 1 || kind == 2, "cast_lp: bad kind");
    long long n = rows * (cols / 4);
    cast_lp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, lds, dst, ldd, rows, cols / 4, kind);
    return check_launch("cast_lp");
}

}  // namespace vog
