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

namespace vog {

constexpr int LS_THREADS = 512;
constexpr int LS_MAXB = 8;          // sequences per launch
constexpr int LS_MAXU = 16;         // hidden units per CTA

struct LstmParams {
    const float* gx; long long ldg;       // [T*Bq, ldg]; direction d at columns d*4H
    const float* whh;                     // [2, 4H, H]
    const long long* lens;                // [Bq]
    float* hbuf;                          // [2 parity, 2 dir, Bq, H]
    unsigned* counters;                   // [2], zeroed before launch
    void* out_lp; long long ld_out; int lp_kind;   // [T*Bq, 2H] bf16 / tf32-rounded fp32
    int T, Bq, H, U, ctas_per_dir;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void store_lp(void* base, long long idx, float v, int kind) {
    if (kind == 1) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(base)[idx] = to_tf32(v);
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
                    v += p.gx[((size_t)t * Bq + lane) * p.ldg + (size_t)d * 4 * H + row_of(ri)];
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
                const float cn = gf * c_s[uu * LS_MAXB + b] + gi * gg;
                c_s[uu * LS_MAXB + b] = cn;
                hval = go * tanhf(cn);
                hc_s[uu * LS_MAXB + b] = hval;
            }
            hb[(size_t)b * H + u0 + uu] = hc_s[uu * LS_MAXB + b];
            store_lp(p.out_lp, ((long long)t * Bq + b) * p.ld_out + (long long)d * H + u0 + uu, hval, p.lp_kind);
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
        store_lp(p.out_lp, ((long long)t * Bq + b) * p.ld_out + (long long)d * H + u0 + uu, 0.f, p.lp_kind);
    }
}

long long lstm_workspace_bytes(int Bq, int H)
{
    return (long long)2 * 2 * Bq * H * 4 + 64;
}

int lstm_layer_fwd(const float* gx, long long ldg, const float* whh, const long long* lens, int T, int Bq,
                   int H, void* out_lp, long long ld_out, int lp_kind, void* workspace, cudaStream_t st)
{
    if (T == 0 || Bq == 0) return 0;
    VOG_REQUIRE(Bq <= LS_MAXB, "lstm_layer_fwd: at most %d sequences per call (got %d)", LS_MAXB, Bq);
    VOG_REQUIRE(H % 4 == 0 && H >= 4 && H <= 1024, "lstm_layer_fwd: H must be a multiple of 4, <= 1024");
    VOG_REQUIRE(lp_kind == 1 || lp_kind == 2, "lstm_layer_fwd: bad lp_kind");
    VOG_REQUIRE(ldg >= 8LL * H && ld_out >= 2LL * H, "lstm_layer_fwd: bad leading dimension");
    VOG_REQUIRE((reinterpret_cast<uintptr_t>(whh) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "lstm_layer_fwd: whh and workspace must be 16-byte aligned");
    int sms = num_sms();
    if (sms < 2) sms = 2;
    int per_dir = sms / 2;
    int U = cdiv(H, per_dir);
    if (U > LS_MAXU) { U = LS_MAXU; }
    per_dir = cdiv(H, U);
    VOG_REQUIRE(2 * per_dir <= sms, "lstm_layer_fwd: H=%d needs %d co-resident CTAs but the device has %d SMs",
                H, 2 * per_dir, sms);
    LstmParams p;
    p.gx = gx; p.ldg = ldg; p.whh = whh; p.lens = lens;
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.hbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 64);
    p.out_lp = out_lp; p.ld_out = ld_out; p.lp_kind = lp_kind;
    p.T = T; p.Bq = Bq; p.H = H; p.U = U; p.ctas_per_dir = per_dir;
    VOG_CUDA(cudaMemsetAsync(workspace, 0, 64, st));
    const size_t smem = sizeof(float) * ((size_t)LS_MAXB * H + 6 * LS_MAXU * LS_MAXB);
    VOG_CUDA(cudaFuncSetAttribute(lstm_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_rec_kernel<<<2 * per_dir, LS_THREADS, smem, st>>>(p);
    return check_launch("lstm_rec");
}

}  // namespace vog
