// LSTM backward through time as ONE persistent launch per layer (nn.LSTM backward of utils/mdl_srl_utils.py:100-152
// under loss.backward(), utils/trn_utils.py:500-505).  Replaces T dependent lstm_bwd_step_kernel launches, each of
// which streamed the 33 MB of W_hh^T from L2 (28 us per step), by the forward recurrence's recipe: the recurrent
// weights stay ON CHIP for the whole sequence and the CTAs exchange only small self-tagged records through L2.
//
//   dh_t   = dout_t + dh_rec_t                          (dh_rec from the step processed before: t+1 forward, t-1 reverse)
//   dc_t   = dc_carry + dh_t * o * (1 - tanh(c_t)^2)
//   dG_t   = [dc*g*i(1-i), dc*c_{t-1}*f(1-f), dc*i*(1-g^2), dh*tanh(c_t)*o(1-o)]          (4H gate pre-activations)
//   dh_rec = W_hh^T dG_t,   dc_carry = dc_t * f
//
// Ownership follows the FORWARD kernel: CTA (direction, c) owns U = 14 hidden units, i.e. the 4U gate rows
// r = (gate, unit) of W_hh.  The gate gradients of those rows need only the CTA's own units (dout, the kept
// activations, the carried dc), so dG never crosses CTAs.  What crosses is the matvec: the CTA multiplies ITS rows
// into a partial sum over all H output columns,
//     P_c[col, b] = sum_{r in rows(c)} W_hh[r, col] * dG_t[r, b],
// (thread = two columns, the 56 row weights of each in registers + conflict-free shared memory, dG rows read as
// shared-memory broadcasts: no reduction inside the CTA), publishes it as one 16 KB record of full 128-byte lines
// (single writer, written once per step, lowest mantissa bit = step tag: see lstm_rec.cu, exchange protocol 2), and
// the owner warp of unit j adds the 74 partial sums P_c'[j, :] - 3 x 16 bytes per lane + a warp all-reduce.
// Per step and CTA: 16 KB written, 16.5 KB read; the weights are read from L2 / HBM once per launch.
#include "common.cuh"
#include "kernels.h"
#include "lstm_xchg.cuh"

namespace vog {

constexpr int LBR_H = 1024;
constexpr int LBR_U = 14;                    // hidden units per CTA (148 SMs: 74 CTAs per direction)
constexpr int LBR_R = 4 * LBR_U;             // gate rows per CTA
constexpr int LBR_THREADS = 512;             // thread t owns output columns t and t + 512
constexpr int LBR_REGROWS = 16;              // rows whose two weights per thread stay in registers
constexpr int LBR_NQ = (LBR_R - LBR_REGROWS) / 2;   // float4 {w0[r], w1[r], w0[r+1], w1[r+1]} per thread in shared memory
constexpr int LBR_BQ = 4;                    // sequences per launch (lanes >= Bq carry zeros)
constexpr int LBR_MAX_CTAS = 128;            // CTAs per direction the workspace is sized for

struct LstmBwdResParams {
    const float* dout;                       // [T*Bq, 2H]
    const float* acts;                       // [T*Bq, 2, 6, H] i, f, g, o, tanh(c_t), c_{t-1}
    const float* whh_t;                      // [2, H, 4H]
    const float* whh;                        // [2, 4H, H] the same weights, not transposed (nullable: coalesced one-time load)
    const long long* lens;                   // [Bq]
    float* dG;                               // [T*Bq, 8H]
    unsigned* xw;                            // [2 parity][2 dir][ctas_per_dir][H][LBR_BQ] self-tagged partial sums
    int T, Bq, ctas_per_dir;
    long long* trace;                        // debug (-DVOG_LSTM_TRACE): [8] clock64 phases of CTA 0 / thread 0
};

__global__ void __launch_bounds__(LBR_THREADS, 1)
lstm_bwd_resident_kernel(const LstmBwdResParams p)
{
    constexpr int H = LBR_H, U = LBR_U, R = LBR_R, BQ = LBR_BQ;
    extern __shared__ __align__(16) float sm[];
    float4* w_s = reinterpret_cast<float4*>(sm);                 // [LBR_NQ][512]
    float* dg_s = sm + LBR_NQ * LBR_THREADS * 4;                 // [2 parity][R][BQ]
    __shared__ int len_s[8];
    __shared__ int tmax_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.x / p.ctas_per_dir, c = blockIdx.x % p.ctas_per_dir;
    const int Bq = p.Bq;
    const int j = c * U + warp;                                  // hidden unit of this warp (warps < U)
    const bool owner = warp < U, unit_ok = owner && j < H;

    // ---- one-time weight load: W_hh[gate*H + c*U + uu, col] = whh_t[dir][col][gate*H + c*U + uu]
    float wr[LBR_REGROWS][2];
    {
        // from W_hh itself when the caller has it: a warp reads 128 contiguous bytes of one gate row per load; from the
        // transposed copy every thread walks its own 16 KB row (56-byte runs: ~30 us of the launch)
        const bool nt = p.whh != nullptr;
        const float* w0 = nt ? p.whh + ((size_t)d * 4 * H + (size_t)c * U) * H + tid
                             : p.whh_t + ((size_t)d * H + tid) * 4 * H + (size_t)c * U;
        const float* w1 = nt ? w0 + LBR_THREADS : w0 + (size_t)LBR_THREADS * 4 * H;
        auto wv = [&](const float* base, int r) {
            const int g = r / U, uu = r % U;
            if (c * U + uu >= H) return 0.f;
            return nt ? __ldg(base + ((size_t)g * H + uu) * H) : __ldg(base + (size_t)g * H + uu);
        };
#pragma unroll
        for (int r = 0; r < LBR_REGROWS; ++r) { wr[r][0] = wv(w0, r); wr[r][1] = wv(w1, r); }
#pragma unroll
        for (int q = 0; q < LBR_NQ; ++q) {
            const int r = LBR_REGROWS + 2 * q;
            w_s[q * LBR_THREADS + tid] = make_float4(wv(w0, r), wv(w1, r), wv(w0, r + 1), wv(w1, r + 1));
        }
    }
    for (int i = tid; i < 2 * R * BQ; i += LBR_THREADS) dg_s[i] = 0.f;
    if (tid < 8) len_s[tid] = tid < Bq ? (int)min((long long)p.T, max(0LL, p.lens[tid])) : 0;
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int b = 0; b < Bq; ++b) m = max(m, len_s[b]);
        tmax_s = m;
    }
    __syncthreads();
    const int Tmax = tmax_s;

    float dh_carry = 0.f, dc_carry = 0.f;                        // lane b of an owner warp: (unit j, sequence b)
#ifdef VOG_LSTM_TRACE
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
    long long tc[5] = {0, 0, 0, 0, 0};
    long long tprev = tr ? clock64() : 0;
#define LBR_TRACE(i) if (tr) { const long long tn = clock64(); tc[i] += tn - tprev; tprev = tn; }
#else
#define LBR_TRACE(i)
#endif
    for (int s = 0; s < Tmax; ++s) {
        const int t = d == 0 ? Tmax - 1 - s : s;                 // reverse of the forward order of this direction
        const int par = s & 1;
        const bool mine = unit_ok && lane < Bq;
        const bool active = mine && t < len_s[lane];
        const long long row = (long long)t * Bq + lane;
        // the step's own inputs do not depend on the recurrence: issue their loads before waiting for the exchange
        float a_i = 0.f, a_f = 0.f, a_g = 0.f, a_o = 0.f, a_tc = 0.f, a_cp = 0.f, dout_v = 0.f;
        if (active) {
            const float* a = p.acts + ((row * 2 + d) * 6) * H + j;
            a_i = __ldg(a); a_f = __ldg(a + H); a_g = __ldg(a + 2 * (size_t)H); a_o = __ldg(a + 3 * (size_t)H);
            a_tc = __ldg(a + 4 * (size_t)H); a_cp = __ldg(a + 5 * (size_t)H);
            dout_v = __ldg(p.dout + row * 2 * H + (size_t)d * H + j);
        }
        // ---- dh_rec[j, :] = sum over the producers' partial sums of the previous step
        if (s > 0 && owner) {
            const unsigned want = ((((unsigned)(s - 1)) >> 1) & 1u) ^ 1u;
            const unsigned* xr = p.xw + ((((size_t)((s - 1) & 1) * 2 + d) * p.ctas_per_dir) * H + (unit_ok ? j : 0)) * BQ;
            float v[BQ];
#pragma unroll
            for (int b = 0; b < BQ; ++b) v[b] = 0.f;
            unsigned pend = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (lane + 32 * k < p.ctas_per_dir) pend |= 1u << k;
            XVec<BQ> w[4];
            long long t0 = 0;
            while (pend) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (pend & (1u << k)) w[k] = ld_relaxed_words<BQ>(xr + (size_t)(lane + 32 * k) * H * BQ);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (pend & (1u << k)) {
                        bool ok = true;
#pragma unroll
                        for (int e = 0; e < BQ; ++e) ok = ok && (w[k].w[e] & 1u) == want;
                        if (ok) pend &= ~(1u << k);                // w[k] is final: summed below in a FIXED order
                    }
                if (pend) {
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 4000000000LL) {
                        printf("vog: lstm backward exchange timeout block %d step %d\n", (int)blockIdx.x, s);
                        __trap();
                    }
                }
            }
            // records arrive in any order; the sum does not depend on it (bit-identical results across launches)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (lane + 32 * k < p.ctas_per_dir) {
#pragma unroll
                    for (int e = 0; e < BQ; ++e) v[e] += __uint_as_float(w[k].w[e] & ~1u);
                }
#pragma unroll
            for (int e = 0; e < BQ; ++e) v[e] = warp_sum(v[e]);
            float mine_v = 0.f;
#pragma unroll
            for (int e = 0; e < BQ; ++e)
                if (e == lane) mine_v = v[e];
            // the carried dh only moves on steps that were live for the sequence (lstm_bwd_step_kernel: dh_out =
            // active ? sum : dh_in); the previous step's t is one further along this direction's backward order
            const int tp = d == 0 ? t + 1 : t - 1;
            if (mine && tp < len_s[lane]) dh_carry = mine_v;
        }
        LBR_TRACE(0)                                               // input loads issued + exchange collected
        // ---- gate gradients of (unit j, sequence lane)
        if (owner) {
            float gi_ = 0.f, gf_ = 0.f, gg_ = 0.f, go_ = 0.f;
            if (active) {
                const float dh = dout_v + dh_carry;
                const float dc = fmaf(dh * a_o, 1.f - a_tc * a_tc, dc_carry);
                go_ = dh * a_tc * a_o * (1.f - a_o);
                gi_ = dc * a_g * a_i * (1.f - a_i);
                gg_ = dc * a_i * (1.f - a_g * a_g);
                gf_ = dc * a_cp * a_f * (1.f - a_f);
                dc_carry = dc * a_f;
            }
            if (mine) {
                float* g = p.dG + row * 8 * H + (size_t)d * 4 * H + j;
                g[0] = gi_; g[H] = gf_; g[2 * (size_t)H] = gg_; g[3 * (size_t)H] = go_;
            }
            if (lane < BQ) {
                float* dg = dg_s + (size_t)par * R * BQ;
                dg[(0 * U + warp) * BQ + lane] = gi_;
                dg[(1 * U + warp) * BQ + lane] = gf_;
                dg[(2 * U + warp) * BQ + lane] = gg_;
                dg[(3 * U + warp) * BQ + lane] = go_;
            }
        }
        LBR_TRACE(1)                                               // gate gradients
        __syncthreads();
        LBR_TRACE(2)                                               // barrier
        // ---- this CTA's rows times dG_t -> partial sums for every output column, published as one record
        if (s + 1 < Tmax) {
            const float4* dg4 = reinterpret_cast<const float4*>(dg_s + (size_t)par * R * BQ);
            float acc0[BQ], acc1[BQ];
#pragma unroll
            for (int b = 0; b < BQ; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
            auto fma_row = [&](int r, float w0, float w1) {
                const float4 g = dg4[r];                           // the same address for every lane: broadcast
                acc0[0] = fmaf(w0, g.x, acc0[0]); acc0[1] = fmaf(w0, g.y, acc0[1]);
                acc0[2] = fmaf(w0, g.z, acc0[2]); acc0[3] = fmaf(w0, g.w, acc0[3]);
                acc1[0] = fmaf(w1, g.x, acc1[0]); acc1[1] = fmaf(w1, g.y, acc1[1]);
                acc1[2] = fmaf(w1, g.z, acc1[2]); acc1[3] = fmaf(w1, g.w, acc1[3]);
            };
#pragma unroll
            for (int r = 0; r < LBR_REGROWS; ++r) fma_row(r, wr[r][0], wr[r][1]);
#pragma unroll
            for (int q = 0; q < LBR_NQ; ++q) {
                const float4 w4 = w_s[q * LBR_THREADS + tid];
                fma_row(LBR_REGROWS + 2 * q, w4.x, w4.y);
                fma_row(LBR_REGROWS + 2 * q + 1, w4.z, w4.w);
            }
            const unsigned tag = (((unsigned)s >> 1) & 1u) ^ 1u;
            unsigned* rec = p.xw + ((((size_t)par * 2 + d) * p.ctas_per_dir + c) * H) * BQ;
            auto put = [&](int col, const float (&a)[BQ]) {
                st_relaxed_u32x4(rec + (size_t)col * BQ, (__float_as_uint(a[0]) & ~1u) | tag, (__float_as_uint(a[1]) & ~1u) | tag,
                                 (__float_as_uint(a[2]) & ~1u) | tag, (__float_as_uint(a[3]) & ~1u) | tag);
            };
            LBR_TRACE(3)                                           // matvec
            put(tid, acc0);
            put(tid + LBR_THREADS, acc1);
            LBR_TRACE(4)                                           // publish
        }
    }
#ifdef VOG_LSTM_TRACE
    if (tr) { for (int i = 0; i < 5; ++i) p.trace[i] = tc[i]; p.trace[5] = Tmax; }
#endif
    // rows past the longest sentence carry no gradient
    if (unit_ok && lane < Bq)
        for (int t = Tmax; t < p.T; ++t) {
            float* g = p.dG + ((long long)t * Bq + lane) * 8 * H + (size_t)d * 4 * H + j;
            g[0] = 0.f; g[H] = 0.f; g[2 * (size_t)H] = 0.f; g[3 * (size_t)H] = 0.f;
        }
}

static thread_local int g_lstm_bwd_resident = 1;
void lstm_bwd_set_resident(int on) { g_lstm_bwd_resident = on ? 1 : 0; }

long long lstm_bwd_workspace_bytes(int Bq, int H)
{
    const long long carry = (long long)8 * Bq * H * 4;                                   // per-step kernels: dh / dc carries
    const long long records = H == LBR_H ? (long long)2 * 2 * LBR_MAX_CTAS * H * LBR_BQ * 4 : 0;
    return carry > records ? carry : records;
}

// -> 1 launched, 0 not applicable (the caller runs the per-step kernels), -1 error
int lstm_bwd_resident(const float* dout, const float* acts, const float* whh_t, const float* whh, const long long* lens, float* dG,
                      void* ws, long long ws_bytes, int T, int Bq, int H, cudaStream_t st)
{
    if (!g_lstm_bwd_resident || H != LBR_H || Bq < 1 || Bq > LBR_BQ) return 0;
    const int sms = num_sms();
    const int per_dir = sms / 2;
    if (per_dir < 1 || cdiv(H, per_dir) != LBR_U) return 0;       // sized for 148 SMs (14 units per CTA)
    const int ctas_per_dir = cdiv(H, LBR_U);
    if (ctas_per_dir > LBR_MAX_CTAS || ctas_per_dir > 128) return 0;
    const long long need = (long long)2 * 2 * ctas_per_dir * H * LBR_BQ * 4;
    if (ws == nullptr || ws_bytes < need) return 0;
    VOG_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0 && (reinterpret_cast<uintptr_t>(whh_t) & 3) == 0,
                "lstm_bwd_steps: workspace must be 16-byte aligned");
    VOG_CUDA(cudaMemsetAsync(ws, 0, (size_t)need, st));           // tags never alias across launches / graph replays
    LstmBwdResParams p;
    p.dout = dout; p.acts = acts; p.whh_t = whh_t; p.whh = whh; p.lens = lens; p.dG = dG;
    p.xw = reinterpret_cast<unsigned*>(ws);
    p.T = T; p.Bq = Bq; p.ctas_per_dir = ctas_per_dir;
    p.trace = lstm_get_trace();
    const size_t smem = (size_t)LBR_NQ * LBR_THREADS * 16 + (size_t)2 * LBR_R * LBR_BQ * 4;
    VOG_CUDA(cudaFuncSetAttribute(lstm_bwd_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_bwd_resident_kernel<<<2 * ctas_per_dir, LBR_THREADS, smem, st>>>(p);
    return check_launch("lstm_bwd_resident") ? -1 : 1;
}

}  // namespace vog
