// Bidirectional LSTM recurrence for the language side (utils/mdl_srl_utils.py:72-175: 2-layer
// bi-LSTM(1024) over the <= 20-token SRL sentence, packed-sequence semantics).
//
// The input projections W_ih x + b_ih + b_hh of ALL timesteps and both directions are one tcgen05
// GEMM ([T*Bq, 512|2048] x [8192, .]^T, vog_tc_gemm); this kernel runs the sequential part
//
//     gates_t = gx_t + W_hh h_{t-1};  c_t = f*c_{t-1} + i*g;  h_t = o*tanh(c_t)      (gate order i,f,g,o)
//
// as ONE persistent launch per layer: every CTA owns a slice of hidden units of one direction for
// the whole sequence (cell state stays in shared memory), streams its fp32 W_hh rows from L2 each
// step (32 MB for both directions: L2-resident after the first step), and the CTAs of a direction
// exchange h_t through a double-buffered global buffer + a monotonically increasing arrival counter
// (software grid barrier; all CTAs are co-resident: grid <= SM count).  Packed-sequence behaviour
// without any host-side length handling: a sequence only updates its state while t < len (so the
// reverse direction starts at its own last token) and emits zeros beyond it - exactly what
// pack_padded_sequence / pad_packed_sequence produce, with no .tolist() synchronisation, which
// makes the whole language side CUDA-graph capturable.
#include "common.cuh"
#include "kernels.h"
#include "lstm_xchg.cuh"

namespace vog {

constexpr int LS_THREADS = 512;
constexpr int LS_MAXB = 8;          // sequences per launch
constexpr int LS_MAXU = 64;         // hidden units per CTA (streaming kernel; 64 = a 32-CTA launch at H = 1024)
constexpr int LS_WS_HEADER = 1024;  // workspace header: barrier counters / per-CTA flags
constexpr int LS_REC_MAX_CTAS = 128;  // record exchange: CTAs per direction the workspace is sized for

struct LstmParams {
    const float* gx; long long ldg;       // [T*Bq, ldg]; direction d at columns d*4H
    const float* whh;                     // [2, 4H, H]
    const long long* lens;                // [Bq]
    float* hbuf;                          // [2 parity, 2 dir, Bq, H]
    unsigned* counters;                   // [2], zeroed before launch
    void* out_lp; long long ld_out; int lp_kind;   // [T*Bq, 2H] bf16 / tf32-rounded fp32
    int T, Bq, H, U, ctas_per_dir;
    int bq_total;                         // sequences per timestep row block of gx / out (>= Bq: chunked batches)
    float* acts;                          // training: [T*bq_total, 2, 6, H] i, f, g, o, tanh(c_t), c_{t-1} (nullable)
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
// tanh(x) = 1 - 2 / (exp(2x) + 1): no cancellation for |x| >~ 0.1, absolute error ~1e-7 near 0
__device__ __forceinline__ float fast_tanh(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

__device__ __forceinline__ void store_lp(void* base, long long idx, float v, int kind) {
    if (kind == 1) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else if (kind == 2) reinterpret_cast<float*>(base)[idx] = to_tf32(v);
    else reinterpret_cast<float*>(base)[idx] = v;          // kind 0: exact fp32 (training keeps the states for the backward)
}

__global__ void __launch_bounds__(LS_THREADS, 1)
lstm_rec_kernel(const LstmParams p)
{
    extern __shared__ float sm[];
    const int H = p.H, Bq = p.Bq, U = p.U;
    float* h_s = sm;                              // [Bq][H]
    float* gate_s = h_s + LS_MAXB * H;            // [4][U][MAXB]
    float* c_s = gate_s + 4 * LS_MAXU * LS_MAXB;  // [U][MAXB]
    float* hc_s = c_s + LS_MAXU * LS_MAXB;        // [U][MAXB] carried hidden state
    __shared__ int len_s[LS_MAXB];
    __shared__ int tmax_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = LS_THREADS / 32;
    const int d = blockIdx.x / p.ctas_per_dir, c = blockIdx.x % p.ctas_per_dir;
    const int u0 = c * U;
    const int nu = max(0, min(U, H - u0));

    if (tid < Bq) len_s[tid] = (int)min((long long)p.T, max(0LL, p.lens[tid]));
    for (int i = tid; i < LS_MAXB * H; i += LS_THREADS) h_s[i] = 0.f;
    for (int i = tid; i < LS_MAXU * LS_MAXB; i += LS_THREADS) { c_s[i] = 0.f; hc_s[i] = 0.f; }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int b = 0; b < Bq; ++b) m = max(m, len_s[b]);
        tmax_s = m;
    }
    __syncthreads();
    const int Tmax = tmax_s;
    const float* whh = p.whh + (size_t)d * 4 * H * H;
    unsigned* ctr = p.counters + d;

    for (int step = 0; step < Tmax; ++step) {
        const int t = d == 0 ? step : Tmax - 1 - step;
        // ---- recurrent matvecs: row ri = gate*nu + uu, one warp per row, all Bq sequences at once
        // the loads of a whole row (H/128 float4 per lane) are issued before any is consumed and the
        // next row's loads are in flight while this one is reduced: the loop is L2-latency bound
        {
            constexpr int NQ = 8;                       // float4 per lane per row: H <= 1024
            const int nrows = 4 * nu;
            float4 wn[NQ];
            auto row_of = [&](int ri) { return (ri / nu) * H + u0 + (ri % nu); };
            auto load_row = [&](int ri, float4 (&w)[NQ]) {
                const float4* w4 = reinterpret_cast<const float4*>(whh + (size_t)row_of(ri) * H);
#pragma unroll
                for (int q = 0; q < NQ; ++q)
                    w[q] = (lane + 32 * q < H / 4) ? __ldg(w4 + lane + 32 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
            };
            if (warp < nrows) load_row(warp, wn);
            for (int ri = warp; ri < nrows; ri += nwarps) {
                float4 w[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) w[q] = wn[q];
                if (ri + nwarps < nrows) load_row(ri + nwarps, wn);
                const int gate = ri / nu, uu = ri % nu;
                float acc[LS_MAXB];
#pragma unroll
                for (int b = 0; b < LS_MAXB; ++b) acc[b] = 0.f;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if (lane + 32 * q < H / 4) {
#pragma unroll
                        for (int b = 0; b < LS_MAXB; ++b) {
                            if (b < Bq) {
                                const float4 hv = *reinterpret_cast<const float4*>(h_s + b * H + 4 * (lane + 32 * q));
                                acc[b] = fmaf(w[q].x, hv.x, acc[b]);
                                acc[b] = fmaf(w[q].y, hv.y, acc[b]);
                                acc[b] = fmaf(w[q].z, hv.z, acc[b]);
                                acc[b] = fmaf(w[q].w, hv.w, acc[b]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < LS_MAXB; ++b)
                    if (b < Bq) acc[b] = warp_sum(acc[b]);
                if (lane < Bq) {
                    float v = 0.f;
#pragma unroll
                    for (int b = 0; b < LS_MAXB; ++b)
                        if (b == lane) v = acc[b];
                    v += p.gx[((size_t)t * p.bq_total + lane) * p.ldg + (size_t)d * 4 * H + row_of(ri)];
                    gate_s[(gate * LS_MAXU + uu) * LS_MAXB + lane] = v;
                }
            }
        }
        __syncthreads();
        // ---- cell update for (unit, sequence) pairs owned by this CTA
        float* hb = p.hbuf + (((size_t)(step & 1) * 2 + d) * Bq) * H;
        if (tid < nu * Bq) {
            const int uu = tid / Bq, b = tid % Bq;
            const bool live = t < len_s[b];
            float hval = 0.f;
            if (live) {
                const float gi = sigmoidf_(gate_s[(0 * LS_MAXU + uu) * LS_MAXB + b]);
                const float gf = sigmoidf_(gate_s[(1 * LS_MAXU + uu) * LS_MAXB + b]);
                const float gg = tanhf(gate_s[(2 * LS_MAXU + uu) * LS_MAXB + b]);
                const float go = sigmoidf_(gate_s[(3 * LS_MAXU + uu) * LS_MAXB + b]);
                const float cp = c_s[uu * LS_MAXB + b];
                const float cn = gf * cp + gi * gg;
                c_s[uu * LS_MAXB + b] = cn;
                const float tcn = tanhf(cn);
                hval = go * tcn;
                hc_s[uu * LS_MAXB + b] = hval;
                if (p.acts != nullptr) {
                    float* a = p.acts + ((((size_t)t * p.bq_total + b) * 2 + d) * 6) * H + u0 + uu;
                    a[0] = gi; a[H] = gf; a[2 * (size_t)H] = gg; a[3 * (size_t)H] = go; a[4 * (size_t)H] = tcn; a[5 * (size_t)H] = cp;
                }
            }
            hb[(size_t)b * H + u0 + uu] = hc_s[uu * LS_MAXB + b];
            store_lp(p.out_lp, ((long long)t * p.bq_total + b) * p.ld_out + (long long)d * H + u0 + uu, hval, p.lp_kind);
        }
        // ---- direction-wide barrier, then fetch the complete h_t (bypassing the non-coherent L1)
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(ctr, 1u);
            const unsigned target = (unsigned)(step + 1) * (unsigned)p.ctas_per_dir;
            long long t0 = clock64();
            while (*reinterpret_cast<volatile unsigned*>(ctr) < target) {
                if (clock64() - t0 > 4000000000LL) {
                    printf("vog: lstm grid barrier timeout block %d step %d\n", (int)blockIdx.x, step);
                    __trap();
                }
            }
            __threadfence();
        }
        __syncthreads();
        if (step + 1 < Tmax) {
            const float4* src = reinterpret_cast<const float4*>(hb);
            for (int i = tid; i < Bq * H / 4; i += LS_THREADS)
                reinterpret_cast<float4*>(h_s)[i] = __ldcg(src + i);
            __syncthreads();
        }
    }
    // rows past the longest sentence: zeros (pad_packed_sequence padding_value=0)
    for (int i = tid; i < (p.T - Tmax) * Bq * nu; i += LS_THREADS) {
        const int uu = i % nu, b = (i / nu) % Bq, t = Tmax + i / (nu * Bq);
        store_lp(p.out_lp, ((long long)t * p.bq_total + b) * p.ld_out + (long long)d * H + u0 + uu, 0.f, p.lp_kind);
    }
}


// -------------------------------------------------------------------------------------------------
// Weight-resident variant (H == 1024, >= 128 SMs): the W_hh rows of a CTA's hidden units never leave
// the SM.  A CTA owns U = ceil(H / ctas_per_dir) units of one direction; warp w owns unit w (its 4
// gate rows), lane l the columns {128 i + 4 l .. +3 : i < 8} of those rows - 128 weights per thread,
// half of them in REGISTERS (i < 4), half in shared memory (i >= 4, float4-interleaved by thread so
// every LDS.128 is conflict free).  2 x 4H x H fp32 = 32 MiB of recurrent weights are thus spread
// over the register files and shared memories of 148 SMs and are read from L2/HBM exactly once per
// launch instead of once per timestep.
//
// Per step a warp computes its 4 x BQ partial dot products, reduce-scatters them over the 32 lanes
// (4*BQ - 1 + log-tail shuffles instead of 5 * 4 * BQ), adds the input projection and runs the cell
// update for its unit in-warp.  h_t is exchanged between the CTAs of a direction WITHOUT a grid
// barrier, through a double-buffered global array that a kernel in front of the launch zeroes (so
// tags never alias across graph replays).  Three protocols (template parameter XM):
//   2 (default) self-tagged RECORDS: a CTA gathers its U x BQ values in shared memory and ONE warp
//     writes them as one record (full 128-byte lines, one writer per line, one write per line and
//     step); the lowest mantissa bit of every fp32 word carries a step-derived tag, so consumers
//     validate each 16-byte load by itself - no flag, no fence, no {value, tag} pairs.  Measured
//     (profiles/r2/lstm_records.txt): the poll + barrier phase went from 9.6 k to 1.2 k cycles per
//     step - with per-unit stores every line of the array was written by 16 warps of 2 CTAs at 16
//     different times while 74 CTAs polled it, and each partial write invalidated and re-fetched
//     the line across the two L2 partitions;
//   0 every value as one 64-bit word {fp32 bits, step tag} stored by its producer lane, consumers
//     poll pairs of words;  1 plain values + one release flag per CTA (kept for A/B).
// -------------------------------------------------------------------------------------------------
constexpr int LR_H = 1024;
constexpr int LR_NI = LR_H / 128;          // float4 per gate row per lane
constexpr int LR_MAXU = 14;                // warps (= hidden units) per CTA
constexpr int LR_THREADS = 32 * LR_MAXU;   // 448 threads
constexpr int LR_PB = 5;                   // 16-byte exchange loads in flight per thread

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {      // (a.x*b.x + c.x, a.y*b.y + c.y), one FFMA2
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}

struct LstmResParams {
    const float* gx; long long ldg;
    const float* whh;
    const long long* lens;
    unsigned long long* hx;               // [2 parity, 2 dir, Bq, H] {value, tag} (xmode 0) / fp32 values (xmode 1) /
                                          // [2 parity, 2 dir, ctas_per_dir, 16 units, BQ] self-tagged fp32 words (xmode 2)
    int backoff_ns;                       // xmode 2: __nanosleep between two polls of a word group that is not there yet
    unsigned* flags;                      // [2 dir * ctas_per_dir] last published step + 1 (xmode 1)
    int xmode;
    void* out_lp; long long ld_out; int lp_kind;
    int T, Bq, U, ctas_per_dir;
    int bq_total;                         // sequences per timestep row block of gx / out (>= Bq: chunked batches)
    long long* trace;                     // debug: [8] accumulated clock64 phases of CTA 0 / thread 0
    float* acts;                          // training: [T*bq_total, 2, 6, H] i, f, g, o, tanh(c_t), c_{t-1} (nullable)
};

// v[0..V) per lane -> lane holds the warp-wide sum of v[idx], idx = the top log2(V) bits of (lane >> (5 - log2 V))
template <int V>
__device__ __forceinline__ float warp_reduce_scatter(float (&v)[V], int lane) {
    static_assert(V == 4 || V == 8 || V == 16 || V == 32, "V must be 4..32");
    int width = 16;
#pragma unroll
    for (int n = V / 2; n >= 1; n >>= 1) {
        const bool up = (lane & width) != 0;
#pragma unroll
        for (int k = 0; k < n; ++k) {
            const float keep = up ? v[k + n] : v[k];
            const float send = up ? v[k] : v[k + n];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, width);
        }
        width >>= 1;
    }
    float r = v[0];
    for (; width >= 1; width >>= 1) r += __shfl_xor_sync(0xffffffffu, r, width);
    return r;
}

// float4 per gate row per lane that stay in registers (the rest lives in shared memory): bounded by the
// 128-register budget of a 448-thread CTA on one side and by 227 KB of shared memory on the other
// (144 registers per thread would hold a third weight group, but 448 threads x 144 registers do not launch:
//  "too many resources requested" - the register file is allocated in larger units than ptxas' 128 suggests)
template <int BQ> struct LrCfg { static constexpr int NREG = BQ <= 4 ? 2 : 3; };
#define LR_BOUNDS __launch_bounds__(LR_THREADS, 1)

template <int BQ, int XM>          // XM: h_t exchange protocol (0 tagged 64-bit words, 1 per-CTA flags, 2 self-tagged records)
__global__ void LR_BOUNDS
lstm_rec_resident_kernel(const LstmResParams p)
{
    constexpr int H = LR_H;
    constexpr int LR_NREG = LrCfg<BQ>::NREG;
    constexpr int V = 4 * BQ;
    extern __shared__ __align__(16) float sm[];
    const int NT = blockDim.x;                               // 32 * U
    float4* w_s = reinterpret_cast<float4*>(sm);             // [4 gates][NI - NREG][NT]
    float* h_s = sm + 4 * (LR_NI - LR_NREG) * NT * 4;        // [2 parity][BQ][H]
    float* gate_s = h_s + 2 * BQ * H;                        // [U][V]
    unsigned* x_s = reinterpret_cast<unsigned*>(gate_s + LR_MAXU * V);   // [U][BQ] this CTA's record of h_t (xmode 2)
    __shared__ int len_s[LS_MAXB];
    __shared__ int tmax_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.x / p.ctas_per_dir, c = blockIdx.x % p.ctas_per_dir;
    const int Bq = p.Bq;
    const int u = c * p.U + warp;                            // hidden unit of this warp
    const bool unit_ok = u < H;

#ifdef VOG_LSTM_TRACE
    const long long t_entry = clock64();
#endif
    pdl_trigger();
    for (int i = tid; i < 2 * BQ * H; i += NT) h_s[i] = 0.f;
    // ---- one-time weight load: row (gate, u), columns 128 i + 4 lane.  W_hh is a parameter (written by plain launches
    //      only), so these 33 MB are fetched BEFORE pdl_wait(): behind the input-projection GEMM that precedes this
    //      kernel in the stream
    float4 wr[4][LR_NREG];
    {
        const float* wbase = p.whh + (size_t)d * 4 * H * H;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4* row = reinterpret_cast<const float4*>(wbase + ((size_t)g * H + (unit_ok ? u : 0)) * H);
#pragma unroll
            for (int i = 0; i < LR_NI; ++i) {
                float4 w4 = unit_ok ? __ldg(row + 32 * i + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < LR_NREG) wr[g][i] = w4;
                else w_s[(g * (LR_NI - LR_NREG) + (i - LR_NREG)) * NT + tid] = w4;
            }
        }
    }
    pdl_wait();                                      // gx, lens and the zeroed exchange workspace are ready
    if (tid < LS_MAXB) len_s[tid] = tid < Bq ? (int)min((long long)p.T, max(0LL, p.lens[tid])) : 0;
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int b = 0; b < Bq; ++b) m = max(m, len_s[b]);
        tmax_s = m;
    }
    __syncthreads();
    const int Tmax = tmax_s;

    // after the reduce-scatter lane l holds output idx = l >> (5 - log2 V), laid out idx = b * 4 + gate
    constexpr int LOGV = V == 4 ? 2 : V == 8 ? 3 : V == 16 ? 4 : 5;
    const int my_idx = lane >> (5 - LOGV);
    const int my_b = my_idx >> 2, my_g = my_idx & 3;
    const bool gx_lane = unit_ok && my_b < Bq && (lane & ((1 << (5 - LOGV)) - 1)) == 0;
    const int npairs = Bq * H / 2;                           // exchange words travel in pairs
    const int NP = (npairs + NT - 1) / NT;                   // pairs per thread
    float c_reg = 0.f, h_reg = 0.f;                          // cell / hidden state of (unit u, sequence lane)
#ifdef VOG_LSTM_TRACE      // build with -DVOG_LSTM_TRACE for the clock64 phase breakdown (profiles/lstm_trace.py)
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
    long long tc[5] = {0, 0, 0, 0, 0};
    long long tprev = tr ? clock64() : 0;
    const long long t_loop = tprev;                          // entry -> first step: zero h, weight load, lens
#define LR_TRACE(i) if (tr) { const long long tn = clock64(); tc[i] += tn - tprev; tprev = tn; }
#else
#define LR_TRACE(i)
#endif

    for (int step = 0; step < Tmax; ++step) {
        const int t = d == 0 ? step : Tmax - 1 - step;
        const float* hcur = h_s + (step & 1) * BQ * H;
        float gxv = 0.f;
        if (gx_lane) gxv = __ldg(p.gx + ((size_t)t * p.bq_total + my_b) * p.ldg + (size_t)d * 4 * H + (size_t)my_g * H + u);
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        if (step > 0) {                                      // h_{-1} = 0
            if constexpr (BQ <= 4) {
                // packed fp32 FMAs (FFMA2): even / odd columns accumulate in the two halves of a register pair
                float2 acc2[V];
#pragma unroll
                for (int k = 0; k < V; ++k) acc2[k] = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < LR_NI; ++i) {
                    float4 w4[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        w4[g] = i < LR_NREG ? wr[g][i < LR_NREG ? i : 0]
                                            : w_s[(g * (LR_NI - LR_NREG) + (i - LR_NREG)) * NT + tid];
#pragma unroll
                    for (int b = 0; b < BQ; ++b) {
                        const float4 hv = *reinterpret_cast<const float4*>(hcur + b * H + 128 * i + 4 * lane);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            acc2[b * 4 + g] = ffma2(make_float2(w4[g].x, w4[g].y), make_float2(hv.x, hv.y), acc2[b * 4 + g]);
                            acc2[b * 4 + g] = ffma2(make_float2(w4[g].z, w4[g].w), make_float2(hv.z, hv.w), acc2[b * 4 + g]);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] = acc2[k].x + acc2[k].y;
            } else {
#pragma unroll
                for (int i = 0; i < LR_NI; ++i) {
                    float4 w4[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        w4[g] = i < LR_NREG ? wr[g][i < LR_NREG ? i : 0]
                                            : w_s[(g * (LR_NI - LR_NREG) + (i - LR_NREG)) * NT + tid];
#pragma unroll
                    for (int b = 0; b < BQ; ++b) {
                        const float4 hv = *reinterpret_cast<const float4*>(hcur + b * H + 128 * i + 4 * lane);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float a = acc[b * 4 + g];
                            a = fmaf(w4[g].x, hv.x, a);
                            a = fmaf(w4[g].y, hv.y, a);
                            a = fmaf(w4[g].z, hv.z, a);
                            a = fmaf(w4[g].w, hv.w, a);
                            acc[b * 4 + g] = a;
                        }
                    }
                }
            }
        }
        LR_TRACE(0)
        const float tot = warp_reduce_scatter<V>(acc, lane);
        if (gx_lane) gate_s[warp * V + my_idx] = tot + gxv;
        __syncwarp();
        LR_TRACE(1)
        // ---- cell update: lane b of the warp owns (unit u, sequence b)
        if (lane < Bq && unit_ok) {
            const bool live = t < len_s[lane];
            float hval = 0.f;
            if (live) {
                // the cell update sits on the per-step critical path of every CTA: sigmoid / tanh through the fast
                // exp (ex2.approx + fast divide, ~2 ulp) instead of the ~40-instruction libm versions
                const float* gs = gate_s + warp * V + lane * 4;
                const float gi = fast_sigmoid(gs[0]);
                const float gf = fast_sigmoid(gs[1]);
                const float gg = fast_tanh(gs[2]);
                const float go = fast_sigmoid(gs[3]);
                const float cp = c_reg;
                c_reg = gf * c_reg + gi * gg;
                const float tcn = fast_tanh(c_reg);
                hval = go * tcn;
                h_reg = hval;
                if (p.acts != nullptr) {             // kept for the backward: exactly the activations this step used
                    float* a = p.acts + ((((size_t)t * p.bq_total + lane) * 2 + d) * 6) * H + u;
                    a[0] = gi; a[H] = gf; a[2 * (size_t)H] = gg; a[3 * (size_t)H] = go; a[4 * (size_t)H] = tcn; a[5 * (size_t)H] = cp;
                }
            }
            if (step + 1 < Tmax && XM != 2) {
                const size_t xi = (((size_t)(step & 1) * 2 + d) * Bq + lane) * H + u;
                if (XM == 0)
                    st_relaxed_u64(p.hx + xi, ((unsigned long long)(unsigned)(step + 1) << 32) | __float_as_uint(h_reg));
                else
                    __stcg(reinterpret_cast<float*>(p.hx) + xi, h_reg);
            }
            store_lp(p.out_lp, ((long long)t * p.bq_total + lane) * p.ld_out + (long long)d * H + u, hval, p.lp_kind);
        }
        __syncwarp();
        // ---- record exchange (xmode 2): the CTA's U x BQ values leave as ONE record written by one warp - every 128-byte
        //      line of the exchange array has a single writer and is written once per step (one coherence event per
        //      line and step instead of one per hidden unit).  Every 32-bit word validates itself: the lowest mantissa
        //      bit carries ((step >> 1) & 1) ^ 1, which differs from what the slot held two steps ago and from the
        //      zeroed array, so no flag, no fence and no 64-bit {value, tag} pairs are needed; the consumer clears the
        //      bit (h_t is used with a 23-bit mantissa whose last bit is zero: |error| <= 2^-24 relative).
        const unsigned tagbit = (((unsigned)step >> 1) & 1u) ^ 1u;
        constexpr int RS = 16 * BQ;                              // words per record (16 unit slots, U <= 14 used)
        unsigned* xrec = reinterpret_cast<unsigned*>(p.hx) + (((size_t)(step & 1) * 2 + d) * p.ctas_per_dir) * RS;
        if (XM == 2 && step + 1 < Tmax) {
            if (lane < BQ) x_s[warp * BQ + lane] = (__float_as_uint(h_reg) & ~1u) | tagbit;   // lanes >= Bq: h_reg == 0
            __syncthreads();
            if (warp == 0) {
                unsigned* mine = xrec + (size_t)c * RS;
                const int nw = (NT >> 5) * BQ;
                for (int i = lane; i < nw; i += 32) st_relaxed_u32(mine + i, x_s[i]);
            }
        }
        LR_TRACE(2)
        if (XM == 2 && step + 1 < Tmax) {
            // ---- collect: word groups of (unit, sequences): item q -> (record q >> 4, unit slot q & 15[, half])
            constexpr int VW = BQ < 4 ? BQ : 4, NV = BQ / VW, MAXI = BQ <= 4 ? 3 : 6;
            const int nitems = p.ctas_per_dir * 16 * NV;
            float* hnext = h_s + ((step + 1) & 1) * BQ * H;
            unsigned pend = 0;
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                const int q = tid + k * NT, slot = q / NV, j = slot & 15;
                if (q < nitems && j < p.U && (slot >> 4) * p.U + j < H) pend |= 1u << k;
            }
            XVec<VW> w[MAXI];
            long long t0 = 0;
            while (pend) {
#pragma unroll
                for (int k = 0; k < MAXI; ++k)
                    if (pend & (1u << k)) {
                        const int q = tid + k * NT, slot = q / NV, part = q % NV;
                        w[k] = ld_relaxed_words<VW>(xrec + (size_t)(slot >> 4) * RS + (slot & 15) * BQ + part * VW);
                    }
#pragma unroll
                for (int k = 0; k < MAXI; ++k)
                    if (pend & (1u << k)) {
                        bool ok = true;
#pragma unroll
                        for (int e = 0; e < VW; ++e) ok = ok && (w[k].w[e] & 1u) == tagbit;
                        if (ok) {
                            pend &= ~(1u << k);
                            const int q = tid + k * NT, slot = q / NV, part = q % NV;
                            const int unit = (slot >> 4) * p.U + (slot & 15);
#pragma unroll
                            for (int e = 0; e < VW; ++e) hnext[(part * VW + e) * H + unit] = __uint_as_float(w[k].w[e] & ~1u);
                        }
                    }
                if (pend) {
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 4000000000LL) {
                        printf("vog: lstm record-exchange timeout block %d step %d\n", (int)blockIdx.x, step);
                        __trap();
                    }
                    if (p.backoff_ns > 0) __nanosleep((unsigned)p.backoff_ns);
                }
            }
            LR_TRACE(3)
            __syncthreads();
            LR_TRACE(4)
        }
        if (XM == 1 && step + 1 < Tmax) {
            // ---- flag exchange: plain fp32 values + one release flag per producer CTA; consumers poll the
            //      <= 74 flags of their direction (little L2 traffic), then read the values once
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                st_relaxed_u32(p.flags + blockIdx.x, (unsigned)(step + 1));
            }
            if (tid < p.ctas_per_dir) {
                const unsigned* f = p.flags + d * p.ctas_per_dir + tid;
                long long t0 = 0;
                while (ld_acquire_u32(f) < (unsigned)(step + 1)) {
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 4000000000LL) {
                        printf("vog: lstm flag timeout block %d step %d\n", (int)blockIdx.x, step);
                        __trap();
                    }
                }
            }
            LR_TRACE(3)
            __syncthreads();
            const float4* src4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.hx) +
                                                                 ((size_t)(step & 1) * 2 + d) * Bq * H);
            float4* hn4 = reinterpret_cast<float4*>(h_s + ((step + 1) & 1) * BQ * H);
            for (int i = tid; i < Bq * H / 4; i += NT) hn4[i] = __ldcg(src4 + i);
            __syncthreads();
            LR_TRACE(4)
        }
        // ---- collect h_t of the whole direction: poll the tagged words
        if (XM == 0 && step + 1 < Tmax) {
            const unsigned long long* src = p.hx + ((size_t)(step & 1) * 2 + d) * Bq * H;
            float* hnext = h_s + ((step + 1) & 1) * BQ * H;
            const unsigned want = (unsigned)(step + 1);
            long long t0 = 0;
            // pairs of adjacent words (16-byte loads; each 64-bit half is single-copy atomic), up to
            // LR_PB pairs in flight per thread: normally the whole share of a thread is one batch
            for (int q0 = 0; q0 < NP; q0 += LR_PB) {
                ulonglong2 w[LR_PB];
                unsigned pend = 0;                            // bit k: pair k of this batch not yet complete
#pragma unroll
                for (int k = 0; k < LR_PB; ++k)
                    if (q0 + k < NP && (q0 + k) * NT + tid < npairs) pend |= 1u << k;
                while (pend) {                                // only the words still missing are re-read
#pragma unroll
                    for (int k = 0; k < LR_PB; ++k)
                        if (pend & (1u << k)) w[k] = ld_relaxed_u64x2(src + 2 * (size_t)((q0 + k) * NT + tid));
#pragma unroll
                    for (int k = 0; k < LR_PB; ++k)
                        if ((pend & (1u << k)) && (unsigned)(w[k].x >> 32) == want && (unsigned)(w[k].y >> 32) == want)
                            pend &= ~(1u << k);
                    if (pend) {
                        if (t0 == 0) t0 = clock64();
                        else if (clock64() - t0 > 4000000000LL) {
                            printf("vog: lstm h-exchange timeout block %d step %d\n", (int)blockIdx.x, step);
                            __trap();
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < LR_PB; ++k) {
                    const int q = (q0 + k) * NT + tid;
                    if (q0 + k < NP && q < npairs)
                        *reinterpret_cast<float2*>(hnext + 2 * q) =
                            make_float2(__uint_as_float((unsigned)w[k].x), __uint_as_float((unsigned)w[k].y));
                }
            }
            LR_TRACE(3)
            __syncthreads();
            LR_TRACE(4)
        }
    }
#ifdef VOG_LSTM_TRACE
    if (tr) { for (int i = 0; i < 5; ++i) p.trace[i] = tc[i]; p.trace[5] = Tmax; p.trace[6] = t_loop - t_entry; }
#endif
    // rows past the longest sentence: zeros (pad_packed_sequence padding_value=0)
    if (lane < Bq && unit_ok)
        for (int t = Tmax; t < p.T; ++t)
            store_lp(p.out_lp, ((long long)t * p.bq_total + lane) * p.ld_out + (long long)d * H + u, 0.f, p.lp_kind);
}

// -------------------------------------------------------------------------------------------------
// Two hidden units per warp (exchange protocol 3 = protocol 2's records, <= 4 sequences per launch).  The matvec of
// the kernel above is bound by shared-memory reads: per step and CTA 172 KB of weights and 229 KB of h_{t-1} - every
// one of the 14 warps reads ALL of h for its single unit.  Here a warp owns two units (8 gate rows): 7 warps read h
// once for two units each (115 KB), the weight traffic stays, the FFMA2 count per SM is unchanged; accumulators
// (8 rows x BQ, as register pairs) and 64 register-resident weights fit the 255-register budget of a 224-thread CTA.
// -------------------------------------------------------------------------------------------------
constexpr int LP_NREG = 2;                     // float4 per gate row per lane kept in registers

template <int BQ>
__global__ void __launch_bounds__(32 * (LR_MAXU / 2), 1)
lstm_rec_pair_kernel(const LstmResParams p)
{
    constexpr int H = LR_H;
    constexpr int V = 8 * BQ;                                // values a warp reduces per step: 2 units x 4 gates x BQ
    constexpr int NS = LR_NI - LP_NREG;                      // float4 per gate row per lane in shared memory
    extern __shared__ __align__(16) float sm[];
    const int NT = blockDim.x;                               // 32 * U / 2
    float4* w_s = reinterpret_cast<float4*>(sm);             // [2 units][4 gates][NS][NT]
    float* h_s = sm + 8 * NS * NT * 4;                       // [2 parity][BQ][H]
    float* gate_s = h_s + 2 * BQ * H;                        // [warps][V]
    unsigned* x_s = reinterpret_cast<unsigned*>(gate_s + (LR_MAXU / 2) * V);   // [16 unit slots][BQ]
    __shared__ int len_s[LS_MAXB];
    __shared__ int tmax_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.x / p.ctas_per_dir, c = blockIdx.x % p.ctas_per_dir;
    const int Bq = p.Bq;
    const int u0 = c * p.U + 2 * warp;                       // the warp's units: u0, u0 + 1

    pdl_trigger();
    for (int i = tid; i < 2 * BQ * H; i += NT) h_s[i] = 0.f;
    float4 wr[2][4][LP_NREG];
    {
        const float* wbase = p.whh + (size_t)d * 4 * H * H;
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
            const bool ok = u0 + uu < H;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const float4* row = reinterpret_cast<const float4*>(wbase + ((size_t)g * H + (ok ? u0 + uu : 0)) * H);
#pragma unroll
                for (int i = 0; i < LR_NI; ++i) {
                    const float4 w4 = ok ? __ldg(row + 32 * i + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < LP_NREG) wr[uu][g][i] = w4;
                    else w_s[((uu * 4 + g) * NS + (i - LP_NREG)) * NT + tid] = w4;
                }
            }
        }
    }
    pdl_wait();
    if (tid < LS_MAXB) len_s[tid] = tid < Bq ? (int)min((long long)p.T, max(0LL, p.lens[tid])) : 0;
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int b = 0; b < Bq; ++b) m = max(m, len_s[b]);
        tmax_s = m;
    }
    __syncthreads();
    const int Tmax = tmax_s;

    // after the reduce-scatter lane l holds output idx = l >> (5 - log2 V), laid out idx = b * 8 + unit * 4 + gate
    constexpr int LOGV = V == 8 ? 3 : V == 16 ? 4 : 5;
    const int my_idx = lane >> (5 - LOGV);
    const int my_b = my_idx >> 3, my_uu = (my_idx >> 2) & 1, my_g = my_idx & 3;
    const bool gx_lane = u0 + my_uu < H && my_b < Bq && (lane & ((1 << (5 - LOGV)) - 1)) == 0;
    // cell update: lane = unit * BQ + sequence
    const int cu = lane / BQ, cb = lane % BQ;
    const bool cell_lane = lane < 2 * BQ && cb < Bq && u0 + cu < H;
    float c_reg = 0.f, h_reg = 0.f;
#ifdef VOG_LSTM_TRACE
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
    long long tc[5] = {0, 0, 0, 0, 0};
    long long tprev = tr ? clock64() : 0;
#define LP_TRACE(i) if (tr) { const long long tn = clock64(); tc[i] += tn - tprev; tprev = tn; }
#else
#define LP_TRACE(i)
#endif

    for (int step = 0; step < Tmax; ++step) {
        const int t = d == 0 ? step : Tmax - 1 - step;
        const float* hcur = h_s + (step & 1) * BQ * H;
        float gxv = 0.f;
        if (gx_lane)
            gxv = __ldg(p.gx + ((size_t)t * p.bq_total + my_b) * p.ldg + (size_t)d * 4 * H + (size_t)my_g * H + u0 + my_uu);
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
        if (step > 0) {                                      // h_{-1} = 0
            float2 acc2[V];
#pragma unroll
            for (int k = 0; k < V; ++k) acc2[k] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < LR_NI; ++i) {
                float4 w4[2][4];
#pragma unroll
                for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        w4[uu][g] = i < LP_NREG ? wr[uu][g][i < LP_NREG ? i : 0]
                                                : w_s[((uu * 4 + g) * NS + (i - LP_NREG)) * NT + tid];
#pragma unroll
                for (int b = 0; b < BQ; ++b) {
                    const float4 hv = *reinterpret_cast<const float4*>(hcur + b * H + 128 * i + 4 * lane);
#pragma unroll
                    for (int uu = 0; uu < 2; ++uu)
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int k = b * 8 + uu * 4 + g;
                            acc2[k] = ffma2(make_float2(w4[uu][g].x, w4[uu][g].y), make_float2(hv.x, hv.y), acc2[k]);
                            acc2[k] = ffma2(make_float2(w4[uu][g].z, w4[uu][g].w), make_float2(hv.z, hv.w), acc2[k]);
                        }
                }
            }
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = acc2[k].x + acc2[k].y;
        }
        LP_TRACE(0)
        const float tot = warp_reduce_scatter<V>(acc, lane);
        if (gx_lane) gate_s[warp * V + my_idx] = tot + gxv;
        __syncwarp();
        LP_TRACE(1)
        if (cell_lane) {
            const int u = u0 + cu;
            const bool live = t < len_s[cb];
            float hval = 0.f;
            if (live) {
                const float* gs = gate_s + warp * V + cb * 8 + cu * 4;
                const float gi = fast_sigmoid(gs[0]);
                const float gf = fast_sigmoid(gs[1]);
                const float gg = fast_tanh(gs[2]);
                const float go = fast_sigmoid(gs[3]);
                const float cp = c_reg;
                c_reg = gf * c_reg + gi * gg;
                const float tcn = fast_tanh(c_reg);
                hval = go * tcn;
                h_reg = hval;
                if (p.acts != nullptr) {
                    float* a = p.acts + ((((size_t)t * p.bq_total + cb) * 2 + d) * 6) * H + u;
                    a[0] = gi; a[H] = gf; a[2 * (size_t)H] = gg; a[3 * (size_t)H] = go; a[4 * (size_t)H] = tcn; a[5 * (size_t)H] = cp;
                }
            }
            store_lp(p.out_lp, ((long long)t * p.bq_total + cb) * p.ld_out + (long long)d * H + u, hval, p.lp_kind);
        }
        __syncwarp();
        // ---- record exchange, as protocol 2 of lstm_rec_resident_kernel (unit slot = 2 * warp + unit of the pair)
        const unsigned tagbit = (((unsigned)step >> 1) & 1u) ^ 1u;
        constexpr int RS = 16 * BQ;
        unsigned* xrec = reinterpret_cast<unsigned*>(p.hx) + (((size_t)(step & 1) * 2 + d) * p.ctas_per_dir) * RS;
        if (step + 1 < Tmax) {
            if (lane < 2 * BQ) x_s[(2 * warp + cu) * BQ + cb] = (__float_as_uint(h_reg) & ~1u) | tagbit;   // idle lanes: h_reg == 0
            __syncthreads();
            if (warp == 0) {
                unsigned* mine = xrec + (size_t)c * RS;
                const int nw = p.U * BQ;
                for (int i = lane; i < nw; i += 32) st_relaxed_u32(mine + i, x_s[i]);
            }
        }
        LP_TRACE(2)
        if (step + 1 < Tmax) {
            constexpr int MAXI = 6;                            // 74 records x 16 slots over 224 threads
            const int nitems = p.ctas_per_dir * 16;
            float* hnext = h_s + ((step + 1) & 1) * BQ * H;
            unsigned pend = 0;
#pragma unroll
            for (int k = 0; k < MAXI; ++k) {
                const int q = tid + k * NT, j = q & 15;
                if (q < nitems && j < p.U && (q >> 4) * p.U + j < H) pend |= 1u << k;
            }
            XVec<BQ> w[MAXI];
            long long t0 = 0;
            while (pend) {
#pragma unroll
                for (int k = 0; k < MAXI; ++k)
                    if (pend & (1u << k)) {
                        const int q = tid + k * NT;
                        w[k] = ld_relaxed_words<BQ>(xrec + (size_t)(q >> 4) * RS + (q & 15) * BQ);
                    }
#pragma unroll
                for (int k = 0; k < MAXI; ++k)
                    if (pend & (1u << k)) {
                        bool ok = true;
#pragma unroll
                        for (int e = 0; e < BQ; ++e) ok = ok && (w[k].w[e] & 1u) == tagbit;
                        if (ok) {
                            pend &= ~(1u << k);
                            const int q = tid + k * NT;
                            const int unit = (q >> 4) * p.U + (q & 15);
#pragma unroll
                            for (int e = 0; e < BQ; ++e) hnext[e * H + unit] = __uint_as_float(w[k].w[e] & ~1u);
                        }
                    }
                if (pend) {
                    if (t0 == 0) t0 = clock64();
                    else if (clock64() - t0 > 4000000000LL) {
                        printf("vog: lstm record-exchange timeout block %d step %d\n", (int)blockIdx.x, step);
                        __trap();
                    }
                }
            }
            LP_TRACE(3)
            __syncthreads();
            LP_TRACE(4)
        }
    }
#ifdef VOG_LSTM_TRACE
    if (tr) { for (int i = 0; i < 5; ++i) p.trace[i] = tc[i]; p.trace[5] = Tmax; p.trace[6] = 0; }
#endif
    if (cell_lane)
        for (int t = Tmax; t < p.T; ++t)
            store_lp(p.out_lp, ((long long)t * p.bq_total + cb) * p.ld_out + (long long)d * H + u0 + cu, 0.f, p.lp_kind);
}

template <int BQ>
static int launch_pair(const LstmResParams& p, int ctas, cudaStream_t st)
{
    const int threads = 32 * (p.U / 2);
    const size_t smem = (size_t)8 * (LR_NI - LP_NREG) * threads * 16 + (size_t)2 * BQ * LR_H * 4 + (size_t)(LR_MAXU / 2) * 8 * BQ * 4 +
                        (size_t)16 * BQ * 4;
    VOG_CUDA(cudaFuncSetAttribute(lstm_rec_pair_kernel<BQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VOG_CUDA(launch_pdl(lstm_rec_pair_kernel<BQ>, dim3(ctas), dim3(threads), smem, st, p));
    return check_launch("lstm_rec_pair");
}

__global__ void __launch_bounds__(256) zero16_kernel(uint4* __restrict__ p, long long n16)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (i < n16) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

template <int BQ, int XM>
static int launch_resident_x(const LstmResParams& p, int ctas, int threads, cudaStream_t st)
{
    const size_t smem = (size_t)4 * (LR_NI - LrCfg<BQ>::NREG) * threads * 16 + (size_t)2 * BQ * LR_H * 4 + (size_t)LR_MAXU * 4 * BQ * 4 +
                        (size_t)16 * BQ * 4;
    VOG_CUDA(cudaFuncSetAttribute(lstm_rec_resident_kernel<BQ, XM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VOG_CUDA(launch_pdl(lstm_rec_resident_kernel<BQ, XM>, dim3(ctas), dim3(threads), smem, st, p));
    return check_launch("lstm_rec_resident");
}

template <int BQ>
static int launch_resident(const LstmResParams& p, int ctas, int threads, cudaStream_t st)
{
    if (p.xmode == 2) return launch_resident_x<BQ, 2>(p, ctas, threads, st);
    if (p.xmode == 1) return launch_resident_x<BQ, 1>(p, ctas, threads, st);
    return launch_resident_x<BQ, 0>(p, ctas, threads, st);
}

static thread_local int g_lstm_force_streaming = 0;
static thread_local int g_lstm_max_ctas = 0;
void lstm_set_max_ctas(int n) { g_lstm_max_ctas = n > 0 ? n : 0; }
static thread_local int g_lstm_xmode = 4;          // default: self-tagged records, two units per warp for 3-4 sequences
                                                   // (profiles/r2/lstm_records.txt)
static thread_local int g_lstm_backoff = 0;
void lstm_set_exchange(int mode)          // bits 0-7: protocol (0 tagged words, 1 flags, 2 records), bits 8-23: poll back-off in ns
{
    g_lstm_xmode = (mode & 0xff) <= 4 ? (mode & 0xff) : 0;        // 3: records + two hidden units per warp; 4: automatic
    g_lstm_backoff = (mode >> 8) & 0xffff;
}
static thread_local long long* g_lstm_trace = nullptr;
void lstm_set_trace(long long* buf) { g_lstm_trace = buf; }
long long* lstm_get_trace() { return g_lstm_trace; }
void lstm_force_streaming(int on) { g_lstm_force_streaming = on; }

long long lstm_workspace_bytes(int Bq, int H)
{
    if (Bq > LS_MAXB) Bq = LS_MAXB;            // larger batches run in groups of LS_MAXB over the same workspace
    const long long tagged = (long long)2 * 2 * Bq * H * 8;                 // {value, tag} words
    const int bqt = Bq <= 1 ? 1 : Bq <= 2 ? 2 : Bq <= 4 ? 4 : 8;            // kernel instantiation
    const long long records = (long long)2 * 2 * LS_REC_MAX_CTAS * 16 * bqt * 4;   // 16-slot records, <= LS_REC_MAX_CTAS per direction
    return (((tagged > records ? tagged : records) + LS_WS_HEADER) + 15) & ~15LL;   // header (counters / flags) + exchange
}

static int lstm_layer_chunk(const float* gx, long long ldg, const float* whh, const long long* lens, int T, int Bq, int bq_total,
                   int H, void* out_lp, long long ld_out, int lp_kind, void* workspace, cudaStream_t st, float* acts)
{
    if (T == 0 || Bq == 0) return 0;
    VOG_REQUIRE(Bq <= LS_MAXB, "lstm_layer_fwd: at most %d sequences per call (got %d)", LS_MAXB, Bq);
    VOG_REQUIRE(H % 4 == 0 && H >= 4 && H <= 1024, "lstm_layer_fwd: H must be a multiple of 4, <= 1024");
    VOG_REQUIRE(lp_kind >= 0 && lp_kind <= 2, "lstm_layer_fwd: bad lp_kind");
    VOG_REQUIRE(ldg >= 8LL * H && ld_out >= 2LL * H, "lstm_layer_fwd: bad leading dimension");
    VOG_REQUIRE((reinterpret_cast<uintptr_t>(whh) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "lstm_layer_fwd: whh and workspace must be 16-byte aligned");
    int sms = num_sms();
    if (sms < 2) sms = 2;
    // SM budget (vog_lstm_set_max_ctas): when the caller runs the recurrence NEXT TO a long independent branch it asks
    // for a launch on a few SMs only - the weight-streaming kernel with many hidden units per CTA - so that the other
    // branch keeps the rest of the GPU (the weight-resident kernel needs every SM's registers and shared memory)
    const bool few = g_lstm_max_ctas > 0 && g_lstm_max_ctas < sms;
    if (few) sms = g_lstm_max_ctas < 2 ? 2 : g_lstm_max_ctas;
    if (!few && H == LR_H && cdiv(H, sms / 2) <= LR_MAXU && !g_lstm_force_streaming) {
        // weight-resident kernel: one warp per hidden unit, <= 16 units per CTA
        const int per_dir_r = sms / 2;
        const int U = cdiv(H, per_dir_r);
        const int per_dir_u = cdiv(H, U);
        LstmResParams rp;
        rp.gx = gx; rp.ldg = ldg; rp.whh = whh; rp.lens = lens;
        rp.hx = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + LS_WS_HEADER);
        rp.flags = reinterpret_cast<unsigned*>(workspace);
        rp.xmode = g_lstm_xmode;
        // automatic: two units per warp pays from 3 sequences on (Bq=4: 104 -> 100 us per layer; Bq=1: 79 -> 85 us, the
        // halved thread count makes the collect phase longer than the matvec gets shorter)
        if (rp.xmode == 4) rp.xmode = (Bq >= 3 && Bq <= 4) ? 3 : 2;
        rp.backoff_ns = g_lstm_backoff;
        {   // the record exchange needs its item count to fit the unrolled poll and the workspace; else tagged words
            const int bqt = Bq <= 1 ? 1 : Bq <= 2 ? 2 : Bq <= 4 ? 4 : 8;
            const int nv = bqt > 4 ? 2 : 1, maxi = bqt <= 4 ? 3 : 6;
            if (rp.xmode == 2 && (per_dir_u > LS_REC_MAX_CTAS || per_dir_u * 16 * nv > maxi * 32 * U)) rp.xmode = 0;
        }
        VOG_REQUIRE(2 * per_dir_u * 4 <= LS_WS_HEADER, "lstm_layer_fwd: too many CTAs for the flag header");
        rp.out_lp = out_lp; rp.ld_out = ld_out; rp.lp_kind = lp_kind;
        rp.T = T; rp.Bq = Bq; rp.U = U; rp.ctas_per_dir = per_dir_u; rp.bq_total = bq_total;
        rp.trace = g_lstm_trace;
        rp.acts = acts;
        // the exchange workspace is zeroed by a KERNEL (not a memset node) so that the chain GEMM -> zero -> recurrence
        // stays kernel-to-kernel and the recurrence can start (and load its weights) behind the GEMM
        {
            const long long n16 = (lstm_workspace_bytes(Bq, H) + 15) / 16;
            VOG_CUDA(launch_pdl(zero16_kernel, dim3((unsigned)((n16 + 255) / 256)), dim3(256), 0, st,
                                reinterpret_cast<uint4*>(workspace), n16));
        }
        const int ctas = 2 * per_dir_u, threads = 32 * U;
        if (rp.xmode == 3) {
            // two units per warp: U even, <= 4 sequences, the poll's item count within its unrolled budget
            if (Bq <= 4 && U % 2 == 0 && U <= LR_MAXU && per_dir_u <= LS_REC_MAX_CTAS && per_dir_u * 16 <= 6 * 32 * (U / 2)) {
                if (Bq == 1) return launch_pair<1>(rp, ctas, st);
                if (Bq == 2) return launch_pair<2>(rp, ctas, st);
                return launch_pair<4>(rp, ctas, st);
            }
            rp.xmode = 2;
        }
        if (Bq == 1) return launch_resident<1>(rp, ctas, threads, st);
        if (Bq == 2) return launch_resident<2>(rp, ctas, threads, st);
        if (Bq <= 4) return launch_resident<4>(rp, ctas, threads, st);
        return launch_resident<8>(rp, ctas, threads, st);
    }
    int per_dir = sms / 2;
    int U = cdiv(H, per_dir);
    if (U > LS_MAXU) { U = LS_MAXU; }
    per_dir = cdiv(H, U);
    VOG_REQUIRE(2 * per_dir <= num_sms(), "lstm_layer_fwd: H=%d needs %d co-resident CTAs but the device has %d SMs",
                H, 2 * per_dir, num_sms());
    VOG_REQUIRE(U * Bq <= LS_THREADS, "lstm_layer_fwd: %d units x %d sequences per CTA exceed the cell-update threads", U, Bq);
    LstmParams p;
    p.gx = gx; p.ldg = ldg; p.whh = whh; p.lens = lens;
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.hbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + LS_WS_HEADER);
    p.out_lp = out_lp; p.ld_out = ld_out; p.lp_kind = lp_kind;
    p.T = T; p.Bq = Bq; p.H = H; p.U = U; p.ctas_per_dir = per_dir; p.bq_total = bq_total;
    p.acts = acts;
    VOG_CUDA(cudaMemsetAsync(workspace, 0, 64, st));
    const size_t smem = sizeof(float) * ((size_t)LS_MAXB * H + 6 * LS_MAXU * LS_MAXB);
    VOG_CUDA(cudaFuncSetAttribute(lstm_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_rec_kernel<<<2 * per_dir, LS_THREADS, smem, st>>>(p);
    return check_launch("lstm_rec");
}

// Batches of more than LS_MAXB sequences run as consecutive launches over groups of LS_MAXB (the recurrences of
// different sentences are independent); gx / out rows stay time-major over the WHOLE batch.
int lstm_layer_fwd(const float* gx, long long ldg, const float* whh, const long long* lens, int T, int Bq,
                   int H, void* out_lp, long long ld_out, int lp_kind, void* workspace, cudaStream_t st, float* acts)
{
    VOG_REQUIRE(lp_kind >= 0 && lp_kind <= 2, "lstm_layer_fwd: bad lp_kind");
    const size_t esz = lp_kind == 1 ? 2 : 4;
    for (int b0 = 0; b0 < Bq; b0 += LS_MAXB) {
        const int nb = Bq - b0 < LS_MAXB ? Bq - b0 : LS_MAXB;
        if (lstm_layer_chunk(gx + (size_t)b0 * ldg, ldg, whh, lens + b0, T, nb, Bq, H,
                             reinterpret_cast<char*>(out_lp) + (size_t)b0 * ld_out * esz, ld_out, lp_kind, workspace, st,
                             acts ? acts + (size_t)b0 * 2 * 6 * H : nullptr))
            return -1;
    }
    return 0;
}

}  // namespace vog
