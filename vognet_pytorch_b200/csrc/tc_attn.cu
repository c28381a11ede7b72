// Fused relative-position-bias attention for sm_100a (tcgen05 + TMA + TMEM), one launch for all
// (sequence, head, 128-query tile) triples:
//
//     O[i,:] = sum_j softmax_j((q_i.k_j + relu(a_i - a_j + b_h)) / sqrt(d_model)) v_j
//
// replaces, per RelEncoderLayer, the reference's python loop over heads with
// bmm -> add bias -> div -> softmax -> bmm on materialised [Bt,N,N] fp32 tensors
// (code/transformer_code.py:136-160,182-186) AND the materialisation of the bias itself
// (code/mdl_vog.py:477-488, utils/mdl_srl_utils.py:30-69): the N x N score / probability / bias
// matrices never exist in HBM.
//
//   warp 0      TMA producer: Q tile once, then rings of K_j and V_j tiles ([64 keys x dhp] each; V is
//               consumed in its natural layout as an MN-major B operand of the PV MMA - no transposed copy)
//               stages; its idle lanes stage the rank-1 bias factor a_j of the 64 keys next to them
//   warp 1      single-thread tcgen05.mma issuer:  S_j = Q K_j^T  (128 x 64, TMEM, double buffered)
//               and O += P_j V_j (128 x dh, TMEM); S_{j+1} is issued before P_j is awaited so the
//               tensor pipe runs while the softmax warps work
//   warps 2-5   one query row per thread: tcgen05.ld S row, bias + scale (exp2 domain), online
//               softmax with LAZY rescaling of the TMEM accumulator (only when a row max grows by
//               more than 2^8), P written as bf16 into a 128B-swizzled smem tile that feeds the PV MMA
//
// Layouts (produced by vog_tc_gemm_qkv): Q,K,V [Bt,H,N,dhp] bf16, head dim zero-padded to dhp
// (multiple of 64); key rows beyond N are zero-filled by TMA.  Output [Bt*N, H*dhp] (bf16, or tf32-rounded fp32), heads
// side by side in padded slots - the A operand of the (column-padded) Wo GEMM.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "philox.cuh"

namespace vog {

using namespace tc;

constexpr int FA_BQ = 128;          // query rows per CTA
constexpr int FA_BKV = 64;          // keys per pipeline stage
constexpr int FA_THREADS = 384;          // warpgroups 0-1: softmax (2 threads per row); warpgroup 2: TMA warp, MMA warp, 2 idle
constexpr int FA_MAX_STAGES = 8;
constexpr float FA_RESCALE_T = 8.0f;   // log2 units

struct AttnParams {
    int Bt, N, H, dhp;
    int dh[VOG_MAX_HEADS];
    float c;                    // log2(e) / sqrt(d_model)
    int bias_mode;              // 0 none, 1 rank-1, 2 dense
    const float* a; int nbox;   // [Bt*nbox, H]
    const float* ak_seq; int ak_ld;   // [Bt*H, ak_ld]: c * a[key % nbox] per key, zero padded to 64
    const float* bpe;           // [H]
    const float* dense;         // [Bt,N,N,H]
    const __nv_bfloat16* q;     // [Bt,H,N,dhp] (v2 kernel: Q rows are read directly, not through TMA)
    void* out; long long ldo; int out_kind;   // 1 bf16, 2 tf32-rounded fp32
    int stages;
    int cluster;                // v2: CTAs (query tiles of one sequence/head) that share multicast K / V loads
    uint32_t tmem_cols;
    long long* prof;            // optional [8] phase cycle counters (block 0, first softmax warp)
    float* lse;                 // optional [Bt*H, N]: log2-domain log-sum-exp of every row (kept for the backward; v2 kernel)
    float drop_p;               // training: dropout on the probabilities (code/transformer_code.py:153), 0 = off
    unsigned long long seed;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(FA_THREADS, 1)
tc_attn_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
               const __grid_constant__ CUtensorMap tma_v, const AttnParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = raw_u32;              // 128B-swizzled tiles need 1024-byte alignment
    uint8_t* smem_gen = smem_raw;
    if ((raw_u32 & 1023u) != 0) { if (threadIdx.x == 0) printf("vog: dynamic smem not 1024-aligned\n"); __trap(); }

    const int q_tile = blockIdx.x, h = blockIdx.y, bt = blockIdx.z;
    const int dhp = p.dhp, N = p.N;
    const int dh = p.dh[h];
    const int nkk = dhp / 64;                         // 64-wide sub-tiles of the head dimension
    const int T = (N + FA_BKV - 1) / FA_BKV;          // key tiles

    // ---- shared memory carve-up (all tile bases 1024-aligned)
    const uint32_t q_bytes = FA_BQ * dhp * 2;
    const uint32_t k_bytes = FA_BKV * dhp * 2;        // nkk sub-tiles of [64 rows x 128 B]
    const uint32_t v_bytes = FA_BKV * dhp * 2;        // nkk sub-tiles of [64 keys x 128 B] (64 head-dim columns each)
    // K and V^T live in SEPARATE rings: a K slot is released as soon as S_j = Q K_j^T has been
    // computed (one whole softmax earlier than the V slot, which PV_j still needs), so the next K
    // tile is requested early enough for its L2 latency to hide behind the softmax of tile j
    const int NS = p.stages;
    const uint32_t q_smem = smem_base;
    const uint32_t p_smem0 = q_smem + q_bytes;        // 2 x [128 x 128 B]
    const uint32_t k_smem0 = p_smem0 + 2 * FA_BQ * 128;
    const uint32_t v_smem0 = k_smem0 + NS * k_bytes;
    const uint32_t bar_off = (v_smem0 - smem_base) + NS * v_bytes;
    const uint32_t bar_base = smem_base + bar_off;
    auto k_full = [&](int s) { return bar_base + 8u * s; };
    auto k_empty = [&](int s) { return bar_base + 8u * (FA_MAX_STAGES + s); };
    auto v_full = [&](int s) { return bar_base + 8u * (2 * FA_MAX_STAGES + s); };
    auto v_empty = [&](int s) { return bar_base + 8u * (3 * FA_MAX_STAGES + s); };
    auto s_full = [&](int b) { return bar_base + 8u * (4 * FA_MAX_STAGES + b); };
    auto p_full = [&](int b) { return bar_base + 8u * (4 * FA_MAX_STAGES + 3 + b); };
    auto p_empty = [&](int b) { return bar_base + 8u * (4 * FA_MAX_STAGES + 5 + b); };
    const uint32_t q_full = bar_base + 8u * (4 * FA_MAX_STAGES + 7);
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_gen + bar_off + 8 * (4 * FA_MAX_STAGES + 8));
    float* xchg = reinterpret_cast<float*>(smem_gen + bar_off + 384);     // [2 parity][128 rows][2 halves]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // roles: warps 0-7 softmax (two warpgroups), warp 8 TMA producer, warp 9 MMA issuer; the control
    // warps carry the HIGHEST warp ids because the SM's issue arbiter favours them
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tma_q);
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_v);
        for (int s = 0; s < NS; ++s) {
            mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1);
            mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1);
        }
        for (int b = 0; b < 3; ++b) mbar_init(s_full(b), 1);
        for (int b = 0; b < 2; ++b) { mbar_init(p_full(b), 256); mbar_init(p_empty(b), 1); }
        mbar_init(q_full, 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_o = tmem_base;                 // columns [0, dhp)
    const uint32_t tmem_s0 = tmem_base + dhp;          // 3 x 64 columns: score tiles are produced two ahead

    const int bh = bt * p.H + h;

    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");      // the control warpgroup hands registers ...
    if (warp == 8) {
        // ================= producer =================
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, q_bytes);
            for (int kk = 0; kk < nkk; ++kk)
                tma_load_3d(q_smem + kk * (FA_BQ * 128), &tma_q, q_full, kk * 64, q_tile * FA_BQ, bh);
        }
        int jk = 0, jv = 0;
        long long t_idle = 0;
        while (jk < T || jv < T) {
            int do_k = 0, do_v = 0;
            if (lane == 0) {
                if (jk < T) do_k = mbar_test_wait(k_empty(jk % NS), (uint32_t)(((jk / NS) & 1) ^ 1));
                if (jv < T) do_v = mbar_test_wait(v_empty(jv % NS), (uint32_t)(((jv / NS) & 1) ^ 1));
                if (!(do_k | do_v)) {                  // bounded polling: a protocol bug traps instead of hanging
                    if (t_idle == 0) t_idle = clock64();
                    else if (clock64() - t_idle > 4000000000LL) { printf("vog: attention producer timeout\n"); __trap(); }
                } else t_idle = 0;
            }
            do_k = __shfl_sync(0xffffffffu, do_k, 0);
            do_v = __shfl_sync(0xffffffffu, do_v, 0);
            if (do_k) {
                if (lane == 0) {
                    const int s = jk % NS;
                    mbar_arrive_expect_tx(k_full(s), k_bytes);
                    for (int kk = 0; kk < nkk; ++kk)
                        tma_load_3d(k_smem0 + s * k_bytes + kk * (FA_BKV * 128), &tma_k, k_full(s), kk * 64,
                                    jk * FA_BKV, bh);
                }
                ++jk;
            }
            if (do_v) {
                const int s = jv % NS;
                if (lane == 0) {
                    mbar_arrive_expect_tx(v_full(s), v_bytes);
                    for (int kk = 0; kk < nkk; ++kk)
                        tma_load_3d(v_smem0 + s * v_bytes + kk * (FA_BKV * 128), &tma_v, v_full(s), kk * 64,
                                    jv * FA_BKV, bh);
                }
                ++jv;
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc(FMT_BF16, FA_BQ, FA_BKV);
            const int n_pv = (dh + 15) & ~15;
            // B = V tile in its natural [key][head-dim] layout = MN-major: 8-key x 128 B swizzle atoms, the next
            // 8 keys 1024 B further (SBO), the next 64 head-dim columns one [64 x 128 B] sub-tile further (LBO)
            const uint32_t idesc_o = umma_idesc(FMT_BF16, FA_BQ, n_pv) | (1u << 16);
            const uint32_t v_lbo = ((uint32_t)(FA_BKV * 128) >> 4) << 16;
            const int ksteps = (dh + 15) / 16;           // skip the all-zero padded tail of the head dim
#ifdef VOG_ATTN_PROFILE
            const bool mprof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
            long long mc[4] = {0, 0, 0, 0};
            long long mt = clock64();
#define VOG_MPROF(i) if (mprof) { const long long tn = clock64(); mc[i] += tn - mt; mt = tn; }
#else
#define VOG_MPROF(i)
#endif
            const uint32_t q_lo = umma_desc_lo(q_smem);
            auto issue_s = [&](int j) {
                const int s = j % NS;
                mbar_wait(k_full(s), (uint32_t)((j / NS) & 1));
                tc_fence_after();
                VOG_MPROF(0)
                const uint32_t b_lo = umma_desc_lo(k_smem0 + s * k_bytes);
                const uint32_t d = tmem_s0 + (j % 3) * FA_BKV;
                umma_bf16_lo<false>(d, q_lo, b_lo, idesc_s);
#pragma unroll
                for (int ks = 1; ks < 16; ++ks) {          // descriptor offsets are immediates
                    if (ks < ksteps)
                        umma_bf16_lo<true>(d, q_lo + (ks >> 2) * (FA_BQ * 128 / 16) + (ks & 3) * 2,
                                           b_lo + (ks >> 2) * (FA_BKV * 128 / 16) + (ks & 3) * 2, idesc_s);
                }
                umma_commit(k_empty(s));
                umma_commit(s_full(j % 3));
                VOG_MPROF(1)
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            if (T > 1) issue_s(1);
            for (int j = 0; j < T; ++j) {
                if (j + 2 < T) issue_s(j + 2);      // buffer (j+2)%3 held S_{j-1}: consumed before P_{j-1} was published
                const int s = j % NS, pb = j & 1;
                mbar_wait(v_full(s), (uint32_t)((j / NS) & 1));
                mbar_wait(p_full(pb), (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                VOG_MPROF(2)
                const uint32_t pa_lo = umma_desc_lo(p_smem0 + pb * (FA_BQ * 128));
                const uint32_t vb_lo = (((v_smem0 + s * v_bytes) >> 4) & 0x3FFF) | v_lbo;
                if (j == 0) umma_bf16_lo<false>(tmem_o, pa_lo, vb_lo, idesc_o);
                else umma_bf16_lo<true>(tmem_o, pa_lo, vb_lo, idesc_o);
#pragma unroll
                for (int k4 = 1; k4 < 4; ++k4)         // 16 keys = two 8-key atoms = 2048 B further per step
                    umma_bf16_lo<true>(tmem_o, pa_lo + 2 * k4, vb_lo + k4 * (2048 >> 4), idesc_o);
                umma_commit(v_empty(s));
                umma_commit(p_empty(pb));
                VOG_MPROF(3)
            }
#ifdef VOG_ATTN_PROFILE
            if (mprof) for (int i = 0; i < 4; ++i) p.prof[8 + i] = mc[i];
#endif
        }
        __syncwarp();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");     // ... to the two softmax warpgroups
        // ================= softmax / correction / epilogue =================
        // two threads per query row: warps 0-3 take key columns [0,32) of every 64-key tile, warps 4-7
        // columns [32,64); the pair (same TMEM lane quarter) exchanges its row maxima through smem and
        // a 64-thread named barrier, keeps PARTIAL row sums (combined once at the end) and splits the
        // accumulator columns between them for the lazy rescale and the epilogue
        const int g = warp & 3;
        const int half = warp >> 2;
        const int row = 32 * g + lane;                  // row inside the tile == TMEM lane
        const int qi = q_tile * FA_BQ + row;            // query index inside the sequence
        const bool row_ok = qi < N;
        const uint32_t lane_addr = (uint32_t)(32 * g) << 16;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + g) : "memory"); };
        float ai = 0.f;
        if (p.bias_mode == 1 && row_ok)
            ai = (__ldg(p.a + ((size_t)bt * p.nbox + qi % p.nbox) * p.H + h) + __ldg(p.bpe + h)) * p.c;
        const float* dense_row = nullptr;
        if (p.bias_mode == 2 && row_ok) dense_row = p.dense + (((size_t)bt * N + qi) * N) * p.H + h;
        float m_run = -1e30f, l_run = 0.f;
        const uint32_t prow = (uint32_t)row * 128u;
        const uint32_t sw = (uint32_t)(row & 7);
        const int ocols = dhp >> 1;                     // accumulator columns owned by this thread's half

#ifdef VOG_ATTN_PROFILE     // build with -DVOG_ATTN_PROFILE for the clock64 phase breakdown (profiles/attn_phases.py)
        const bool do_prof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 0 && lane == 0;
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = do_prof ? clock64() : 0;
#define VOG_PROF(i) if (do_prof) { const long long tn = clock64(); pc[i] += tn - tprev; tprev = tn; }
#else
#define VOG_PROF(i)
#endif
        // software pipeline: the score tile S_{j+1} (TMEM) and its bias factors (global, L1/L2) are
        // requested while tile j is exponentiated / packed, so their latencies are off the critical path
        uint32_t rn[32];
        float4 an[8];
        auto load_bias = [&](int jj) {
            const float4* ak4 = reinterpret_cast<const float4*>(p.ak_seq + (size_t)bh * p.ak_ld +
                                                                (size_t)jj * FA_BKV + half * 32);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) an[c4] = __ldg(ak4 + c4);
        };
        auto load_scores = [&](int jj) {
            mbar_wait(s_full(jj % 3), (uint32_t)((jj / 3) & 1));
            tc_fence_after();
            tmem_ld32(tmem_s0 + lane_addr + (jj % 3) * FA_BKV + half * 32, rn);
        };
        if (p.bias_mode == 1) load_bias(0);
        load_scores(0);
        for (int j = 0; j < T; ++j) {
            const int sb = j & 1;
            tmem_wait_ld();
            VOG_PROF(0)
            uint32_t (&r0)[32] = rn;
            float4 (&a4)[8] = an;
            VOG_PROF(1)
            float sv[32];
            if (p.bias_mode == 1) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    sv[4 * c4 + 0] = fmaf(__uint_as_float(r0[4 * c4 + 0]), p.c, fmaxf(ai - a4[c4].x, 0.f));
                    sv[4 * c4 + 1] = fmaf(__uint_as_float(r0[4 * c4 + 1]), p.c, fmaxf(ai - a4[c4].y, 0.f));
                    sv[4 * c4 + 2] = fmaf(__uint_as_float(r0[4 * c4 + 2]), p.c, fmaxf(ai - a4[c4].z, 0.f));
                    sv[4 * c4 + 3] = fmaf(__uint_as_float(r0[4 * c4 + 3]), p.c, fmaxf(ai - a4[c4].w, 0.f));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) sv[c] = __uint_as_float(r0[c]) * p.c;
                if (p.bias_mode == 2 && row_ok) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int key = j * FA_BKV + half * 32 + c;
                        if (key < N) sv[c] += __ldg(dense_row + (size_t)key * p.H) * p.c;
                    }
                }
            }
            if ((j + 1) * FA_BKV > N) {                // ragged last tile: keys >= N do not exist
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    if (j * FA_BKV + half * 32 + c >= N) sv[c] = -INFINITY;
            }
            float mx0 = fmaxf(sv[0], sv[1]), mx1 = fmaxf(sv[2], sv[3]), mx2 = fmaxf(sv[4], sv[5]), mx3 = fmaxf(sv[6], sv[7]);
#pragma unroll
            for (int c = 8; c < 32; c += 4) {
                mx0 = fmaxf(mx0, sv[c]); mx1 = fmaxf(mx1, sv[c + 1]); mx2 = fmaxf(mx2, sv[c + 2]); mx3 = fmaxf(mx3, sv[c + 3]);
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));     // -inf when this half has no valid key
            float* xr = xchg + ((j & 1) * FA_BQ + row) * 2;
            VOG_PROF(2)
            xr[half] = mx;
            pair_sync();
            VOG_PROF(3)
            const float m_new = fmaxf(m_run, fmaxf(mx, xr[half ^ 1]));
            if (j == 0) {
                m_run = m_new;
            } else {
                const bool grow = (m_new - m_run) > FA_RESCALE_T;
                if (__any_sync(0xffffffffu, grow)) {
                    // rescale this thread's half of the accumulator row; PV_{j-1} must have landed first
                    const float alpha = grow ? fast_exp2(m_run - m_new) : 1.f;
                    if (grow) { m_run = m_new; l_run *= alpha; }
                    mbar_wait(p_empty((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));
                    tc_fence_after();
                    for (int c0 = half * ocols; c0 < (half + 1) * ocols; c0 += 32) {
                        uint32_t o[32];
                        tmem_ld32(tmem_o + lane_addr + c0, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
                        tmem_st32(tmem_o + lane_addr + c0, o);
                    }
                    tmem_wait_st();
                }
            }
            if (j + 1 < T) {                           // sv holds tile j now: rn / an are free again
                if (p.bias_mode == 1) load_bias(j + 1);
                load_scores(j + 1);
            }
            VOG_PROF(7)
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                sv[c] = fast_exp2(sv[c] - m_run); s0 += sv[c];
                sv[c + 1] = fast_exp2(sv[c + 1] - m_run); s1 += sv[c + 1];
                sv[c + 2] = fast_exp2(sv[c + 2] - m_run); s2 += sv[c + 2];
                sv[c + 3] = fast_exp2(sv[c + 3] - m_run); s3 += sv[c + 3];
            }
            l_run += (s0 + s1) + (s2 + s3);
            VOG_PROF(4)
            // P buffer (j&1) was last read by PV_{j-2}
            if (j >= 2) mbar_wait(p_empty(sb), (uint32_t)(((j - 2) >> 1) & 1));
            VOG_PROF(5)
            const uint32_t pdst = p_smem0 + sb * (FA_BQ * 128) + prow;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t c16 = (uint32_t)(half * 4 + i);
                const uint32_t w0 = pack_bf16(sv[8 * i + 0], sv[8 * i + 1]);
                const uint32_t w1 = pack_bf16(sv[8 * i + 2], sv[8 * i + 3]);
                const uint32_t w2 = pack_bf16(sv[8 * i + 4], sv[8 * i + 5]);
                const uint32_t w3 = pack_bf16(sv[8 * i + 6], sv[8 * i + 7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(pdst + ((c16 ^ sw) << 4)), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full(sb));
            VOG_PROF(6)
        }
#ifdef VOG_ATTN_PROFILE
        if (do_prof) { for (int i = 0; i < 8; ++i) p.prof[i] = pc[i]; p.prof[12] = T; }
#endif
        // ---- epilogue: O / l  (row sum = the two partial sums of the pair)
        {
            float* xr = xchg + ((T & 1) * FA_BQ + row) * 2;
            xr[half] = l_run;
            pair_sync();
            l_run += xr[half ^ 1];
        }
        mbar_wait(p_empty((T - 1) & 1), (uint32_t)(((T - 1) >> 1) & 1));
        tc_fence_after();
        const float inv_l = 1.f / l_run;
        const int n_pv = (dh + 15) & ~15;
        for (int c0 = half * ocols; c0 < (half + 1) * ocols; c0 += 32) {
            uint32_t o[32];
            float v[32];
            if (c0 < n_pv) {
                tmem_ld32(tmem_o + lane_addr + c0, o);
                tmem_wait_ld();
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = (c0 + c < n_pv) ? __uint_as_float(o[c]) * inv_l : 0.f;
            if (row_ok) {
                const size_t off = ((size_t)bt * N + qi) * p.ldo + (size_t)h * dhp + c0;
                if (p.out_kind == 1) {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
#pragma unroll
                    for (int c = 0; c < 32; c += 8)
                        *reinterpret_cast<uint4*>(dst + c) =
                            make_uint4(pack_bf16(v[c], v[c + 1]), pack_bf16(v[c + 2], v[c + 3]),
                                       pack_bf16(v[c + 4], v[c + 5]), pack_bf16(v[c + 6], v[c + 7]));
                } else {
                    float* dst = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
                    for (int c = 0; c < 32; c += 4)
                        *reinterpret_cast<float4*>(dst + c) =
                            make_float4(to_tf32(v[c]), to_tf32(v[c + 1]), to_tf32(v[c + 2]), to_tf32(v[c + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, p.tmem_cols);
}

// =================================================================================================
// v2: Q and P in TENSOR MEMORY.  The v1 kernel above re-reads the 128 x dhp Q tile from shared memory
// for every 64-key score MMA (64 KB per tile at dhp = 256: the score MMAs are shared-memory-bandwidth
// bound) and round-trips P through shared memory.  Here
//   * the softmax warps load the CTA's Q rows straight from global memory once and park them in TMEM
//     (128 columns of bf16 pairs); S = Q K^T is issued in the TS form (A from TMEM, B = K tile via smem
//     descriptor), so a tile's smem traffic is just the K and V tiles;
//   * P_j (bf16 pairs, 32 columns) overwrites S_j in place with tcgen05.st and feeds PV_j as a TMEM A
//     operand - no st.shared / fence.proxy.async / P buffers;
//   * the freed 96 KB of shared memory turn the K / V rings from 2 into 3+ stages.
// S is double buffered; the issue order S_0 S_1 | PV_0 S_2 | PV_1 S_3 ... makes the in-order tensor pipe
// resolve the write-after-read on the buffer P_j was read from.  TMEM: O [0,256) | Q [256,384) | S/P 2 x 64.
// =================================================================================================
// NPART = softmax threads per query row (2 or 4): 4*NPART softmax warps (warp w: TMEM lane quarter w % 4, key
// columns [(w / 4) * 64/NPART, ...) of every tile) + one control warpgroup (K producer, MMA issuer, V producer).
// With 4 threads per row every scheduler holds four softmax warps instead of two - the softmax is latency /
// issue bound, not MUFU bound - at 104 registers per softmax thread.
template <int NPART>
__global__ void __launch_bounds__(128 * NPART + 128, 1)
tc_attn2_kernel(const __grid_constant__ CUtensorMap tma_k, const __grid_constant__ CUtensorMap tma_v,
                const AttnParams p)
{
    constexpr int CW = 4 * NPART;                     // first control warp
    constexpr int EL = FA_BKV / NPART;                // keys of a tile handled by one softmax thread
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = raw_u32;              // 128B-swizzled tiles need 1024-byte alignment
    uint8_t* smem_gen = smem_raw;
    if ((raw_u32 & 1023u) != 0) { if (threadIdx.x == 0) printf("vog: dynamic smem not 1024-aligned\n"); __trap(); }

    const int q_tile = blockIdx.x, h = blockIdx.y, bt = blockIdx.z;
    const int dhp = p.dhp, N = p.N;
    const int dh = p.dh[h];
    const int nkk = dhp / 64;                         // 64-wide sub-tiles of the head dimension
    const int T = (N + FA_BKV - 1) / FA_BKV;          // key tiles

    // ---- shared memory: ONLY the K and V rings (Q and P live in tensor memory)
    const uint32_t k_bytes = FA_BKV * dhp * 2;        // nkk sub-tiles of [64 rows x 128 B]
    const uint32_t v_bytes = FA_BKV * dhp * 2;        // nkk sub-tiles of [64 keys x 128 B] (64 head-dim columns each)
    const int NS = p.stages;
    const uint32_t k_smem0 = smem_base;
    const uint32_t v_smem0 = k_smem0 + NS * k_bytes;
    const uint32_t bar_off = (v_smem0 - smem_base) + NS * v_bytes;
    const uint32_t bar_base = smem_base + bar_off;
    auto k_full = [&](int s) { return bar_base + 8u * s; };
    auto k_empty = [&](int s) { return bar_base + 8u * (FA_MAX_STAGES + s); };
    auto v_full = [&](int s) { return bar_base + 8u * (2 * FA_MAX_STAGES + s); };
    auto v_empty = [&](int s) { return bar_base + 8u * (3 * FA_MAX_STAGES + s); };
    auto s_full = [&](int b) { return bar_base + 8u * (4 * FA_MAX_STAGES + b); };        // S_j landed in buffer b
    auto p_full = [&](int b) { return bar_base + 8u * (4 * FA_MAX_STAGES + 2 + b); };    // P_j written over S_j
    auto pv_done = [&](int b) { return bar_base + 8u * (4 * FA_MAX_STAGES + 4 + b); };   // PV_j accumulated into O
    const uint32_t q_ready = bar_base + 8u * (4 * FA_MAX_STAGES + 6);
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_gen + bar_off + 8 * (4 * FA_MAX_STAGES + 8));
    float* xchg = reinterpret_cast<float*>(smem_gen + bar_off + 384);     // [2 parity][128 rows][NPART]

    // thread-block cluster of C CTAs = C query tiles of the same (sequence, head): every K / V tile is fetched
    // from L2 ONCE per cluster - CTA r loads the 64/C-key slice r and multicasts it into all C shared memories
    const int C = p.cluster;
    const uint32_t crank = C > 1 ? cluster_ctarank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << C) - 1u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // roles: warps 0-7 softmax (two warpgroups), warp 8 TMA producer, warp 9 MMA issuer; the control
    // warps carry the HIGHEST warp ids because the SM's issue arbiter favours them
    if (warp == CW && lane == 0) {
        tma_prefetch_desc(&tma_k);
        tma_prefetch_desc(&tma_v);
        for (int s = 0; s < NS; ++s) {
            mbar_init(k_full(s), 1); mbar_init(k_empty(s), C);        // a slot is free when ALL CTAs of the cluster read it
            mbar_init(v_full(s), 1); mbar_init(v_empty(s), C);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(p_full(b), 128 * NPART); mbar_init(pv_done(b), 1); }
        mbar_init(q_ready, 256);
        fence_barrier_init();
    }
    if (warp == CW + 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();                     // peers' barriers are initialised before anyone multicasts
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_trigger();
    pdl_wait();                                        // set-up above overlapped the previous kernel's tail
    const uint32_t tmem_o = tmem_base;                 // columns [0, 256): O accumulator (dhp <= 256 used)
    const uint32_t tmem_q = tmem_base + 256;           // columns [256, 384): Q tile, bf16 pairs (A operand of S = Q K^T)
    const uint32_t tmem_s0 = tmem_base + 384;          // 2 x 64 columns: S_j (fp32), overwritten in place by P_j (bf16 pairs)

    const int bh = bt * p.H + h;

    if (warp >= CW) {
    if constexpr (NPART == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");      // the control warpgroup hands registers ...
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == CW || warp == CW + 2) {
        // ================= producers: warp 8 streams the K tiles, warp 10 the V tiles =================
        // Warp-uniform loops with BLOCKING mbarrier waits (a waiting warp takes no issue slots from the softmax
        // warps of its scheduler) and one elected lane per TMA instruction.  The former single polling lane
        // under `if (lane == 0)` issued every K tile ~5000 cycles late (profiles/r1/attention.md): each
        // UTMALDG sat in an ELECT / BRA.U.ANY loop and the poll loop starved next to two busy softmax warps.
        const bool is_k = warp == CW;
        const CUtensorMap* tm = is_k ? &tma_k : &tma_v;
        const uint32_t ring0 = is_k ? k_smem0 : v_smem0;
        for (int j = 0; j < T; ++j) {
            const int s = j % NS;
            const uint32_t full = is_k ? k_full(s) : v_full(s);
            mbar_wait(is_k ? k_empty(s) : v_empty(s), (uint32_t)(((j / NS) & 1) ^ 1));
            if (elect_one()) {
#ifdef VOG_ATTN_PROFILE
                if (is_k && p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 32)
                    p.prof[32 + j] = clock64();                   // K_j requested
#endif
                mbar_arrive_expect_tx(full, k_bytes);
                if (C == 1) {
                    for (int kk = 0; kk < nkk; ++kk)
                        tma_load_3d(ring0 + s * k_bytes + kk * (FA_BKV * 128), tm, full, kk * 64, j * FA_BKV, bh);
                } else {
                    const int rows = FA_BKV / C;                   // this CTA's key slice of the tile
                    for (int kk = 0; kk < nkk; ++kk)
                        tma_load_3d_mc(ring0 + s * k_bytes + kk * (FA_BKV * 128) + crank * rows * 128, tm, full,
                                       kk * 64, j * FA_BKV + crank * rows, bh, cmask);
                }
            }
            __syncwarp();
        }
    } else if (warp == CW + 1) {
        // ================= MMA issuer =================
        // The WHOLE warp runs the control flow (waits, loop counters stay warp-uniform) and one elected lane
        // issues the tcgen05 instructions: under a divergent `if (lane == 0)` the compiler has to wrap every
        // uniform-datapath instruction (UTCHMMA, UTCBAR) in an ELECT / BRA.U.ANY loop, ~11 SASS instructions
        // and ~64 cycles per MMA - twice the 32-cycle dispatch floor of a 128x64x16 MMA.
        {
            const uint32_t idesc_s = umma_idesc(FMT_BF16, FA_BQ, FA_BKV);
            const int n_pv = (dh + 15) & ~15;
            // B = V tile in its natural [key][head-dim] layout = MN-major: 8-key x 128 B swizzle atoms, the next
            // 8 keys 1024 B further (SBO), the next 64 head-dim columns one [64 x 128 B] sub-tile further (LBO)
            const uint32_t idesc_o = umma_idesc(FMT_BF16, FA_BQ, n_pv) | (1u << 16);
            const uint32_t v_lbo = ((uint32_t)(FA_BKV * 128) >> 4) << 16;
            const int ksteps = (dh + 15) / 16;           // skip the all-zero padded tail of the head dim
#ifdef VOG_ATTN_PROFILE
            const bool mprof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
            long long mc[5] = {0, 0, 0, 0, 0};
            long long mt = clock64();
#undef VOG_MPROF
#define VOG_MPROF(i) if (mprof) { const long long tn = clock64(); mc[i] += tn - mt; mt = tn; }
#else
#undef VOG_MPROF
#define VOG_MPROF(i)
#endif
            auto issue_s = [&](int j) {
                const int s = j % NS;
#ifdef VOG_ATTN_PROFILE
                if (mprof && j < 32) p.prof[64 + j] = clock64();  // MMA thread starts waiting for K_j
#endif
                mbar_wait(k_full(s), (uint32_t)((j / NS) & 1));
                tc_fence_after();
#ifdef VOG_ATTN_PROFILE
                if (mprof && j < 32) p.prof[96 + j] = clock64();  // K_j observed in shared memory
#endif
                VOG_MPROF(0)
                const uint32_t b_lo = umma_desc_lo(k_smem0 + s * k_bytes);
                const uint32_t d = tmem_s0 + (j & 1) * FA_BKV;
                if (elect_one()) {
                    umma_bf16_ts<false>(d, tmem_q, b_lo, idesc_s);
#pragma unroll
                    for (int ks = 1; ks < 16; ++ks) {      // A: 16 bf16 = 8 TMEM columns per step; B offsets are immediates
                        if (ks < ksteps)
                            umma_bf16_ts<true>(d, tmem_q + ks * 8,
                                               b_lo + (ks >> 2) * (FA_BKV * 128 / 16) + (ks & 3) * 2, idesc_s);
                    }
                    if (C == 1) umma_commit(k_empty(s)); else umma_commit_mc(k_empty(s), cmask);
                    umma_commit(s_full(j & 1));
                }
                __syncwarp();
                VOG_MPROF(1)
            };
            mbar_wait(q_ready, 0);
            tc_fence_after();
            issue_s(0);
            if (T > 1) issue_s(1);
            for (int j = 0; j < T; ++j) {
                const int s = j % NS, pb = j & 1;
                mbar_wait(v_full(s), (uint32_t)((j / NS) & 1));
                VOG_MPROF(4)
                mbar_wait(p_full(pb), (uint32_t)((j >> 1) & 1));
                tc_fence_after();
                VOG_MPROF(2)
                const uint32_t pa = tmem_s0 + pb * FA_BKV;         // P_j: 64 keys = 32 columns of bf16 pairs
                const uint32_t vb_lo = (((v_smem0 + s * v_bytes) >> 4) & 0x3FFF) | v_lbo;
                if (elect_one()) {
                    if (j == 0) umma_bf16_ts<false>(tmem_o, pa, vb_lo, idesc_o);
                    else umma_bf16_ts<true>(tmem_o, pa, vb_lo, idesc_o);
#pragma unroll
                    for (int k4 = 1; k4 < 4; ++k4)     // 16 keys = 8 P columns / two 8-key V atoms = 2048 B per step
                        umma_bf16_ts<true>(tmem_o, pa + 8 * k4, vb_lo + k4 * (2048 >> 4), idesc_o);
                    if (C == 1) umma_commit(v_empty(s)); else umma_commit_mc(v_empty(s), cmask);
                    umma_commit(pv_done(pb));
                }
                __syncwarp();
                VOG_MPROF(3)
                // S_{j+2} reuses the buffer P_j was read from: issued after PV_j, and the tensor pipe runs in order
                if (j + 2 < T) issue_s(j + 2);
            }
#ifdef VOG_ATTN_PROFILE
            if (mprof) { for (int i = 0; i < 4; ++i) p.prof[8 + i] = mc[i]; p.prof[13] = mc[4]; }
#endif
        }
    }
    } else {
        if constexpr (NPART == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");     // ... to the softmax warpgroups
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ================= softmax / correction / epilogue =================
        // NPART threads per query row: warp w takes key columns [part*EL, part*EL + EL) of every 64-key tile,
        // part = w / 4; the threads of a row (same TMEM lane quarter) exchange their row maxima through smem
        // and a named barrier, keep PARTIAL row sums (combined once at the end) and split the accumulator
        // columns between them for the lazy rescale and the epilogue
        const int g = warp & 3;
        const int part = warp >> 2;
        const int row = 32 * g + lane;                  // row inside the tile == TMEM lane
        const int qi = q_tile * FA_BQ + row;            // query index inside the sequence
        const bool row_ok = qi < N;
        const uint32_t lane_addr = (uint32_t)(32 * g) << 16;
        auto row_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(32 * NPART) : "memory"); };
        float ai = 0.f;
        if (p.bias_mode == 1 && row_ok)
            ai = (__ldg(p.a + ((size_t)bt * p.nbox + qi % p.nbox) * p.H + h) + __ldg(p.bpe + h)) * p.c;
        const float* dense_row = nullptr;
        if (p.bias_mode == 2 && row_ok) dense_row = p.dense + (((size_t)bt * N + qi) * N) * p.H + h;
        float m_run = -1e30f, l_run = 0.f;
        const int ocols = dhp / NPART;                  // accumulator columns owned by this thread (multiple of 16)

#ifdef VOG_ATTN_PROFILE     // build with -DVOG_ATTN_PROFILE for the clock64 phase breakdown (profiles/attn_phases.py)
        const bool do_prof = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && warp == 0 && lane == 0;
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tprev = do_prof ? clock64() : 0;
#undef VOG_PROF
#define VOG_PROF(i) if (do_prof) { const long long tn = clock64(); pc[i] += tn - tprev; tprev = tn; }
#else
#undef VOG_PROF
#define VOG_PROF(i)
#endif
        // ---- Q tile -> tensor memory (A operand of every S = Q K^T of this CTA: never re-read from smem): done by
        //      the first two threads of every row, each one head-dim half = dhp/4 packed words at columns part*dhp/4
        if (part < 2) {
            const uint4* qsrc = reinterpret_cast<const uint4*>(
                p.q + (((size_t)bh * N + (row_ok ? qi : 0)) * dhp) + part * (dhp >> 1));
            for (int c16 = 0; c16 < (dhp >> 6); ++c16) {           // 16 words (32 bf16) per tcgen05.st
                uint32_t w[16];
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const uint4 t = row_ok ? __ldg(qsrc + c16 * 4 + v4) : make_uint4(0u, 0u, 0u, 0u);
                    w[4 * v4] = t.x; w[4 * v4 + 1] = t.y; w[4 * v4 + 2] = t.z; w[4 * v4 + 3] = t.w;
                }
                tmem_st16(tmem_q + lane_addr + part * (dhp >> 2) + c16 * 16, w);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(q_ready);
        }
        uint32_t rn[EL];
        float4 an[EL / 4];
        auto load_bias = [&](int jj) {
            const float4* ak4 = reinterpret_cast<const float4*>(p.ak_seq + (size_t)bh * p.ak_ld +
                                                                (size_t)jj * FA_BKV + part * EL);
#pragma unroll
            for (int c4 = 0; c4 < EL / 4; ++c4) an[c4] = __ldg(ak4 + c4);
        };
        auto load_scores = [&](int jj) {
            mbar_wait(s_full(jj & 1), (uint32_t)((jj >> 1) & 1));
            tc_fence_after();
            if constexpr (EL == 32) tmem_ld32(tmem_s0 + lane_addr + (jj & 1) * FA_BKV + part * EL, rn);
            else tmem_ld16(tmem_s0 + lane_addr + (jj & 1) * FA_BKV + part * EL, rn);
        };
        if (p.bias_mode == 1) load_bias(0);
        load_scores(0);
        for (int j = 0; j < T; ++j) {
            const int sb = j & 1;
            tmem_wait_ld();
            VOG_PROF(0)
            VOG_PROF(1)
            float sv[EL];
            if (p.bias_mode == 1) {
#pragma unroll
                for (int c4 = 0; c4 < EL / 4; ++c4) {
                    sv[4 * c4 + 0] = fmaf(__uint_as_float(rn[4 * c4 + 0]), p.c, fmaxf(ai - an[c4].x, 0.f));
                    sv[4 * c4 + 1] = fmaf(__uint_as_float(rn[4 * c4 + 1]), p.c, fmaxf(ai - an[c4].y, 0.f));
                    sv[4 * c4 + 2] = fmaf(__uint_as_float(rn[4 * c4 + 2]), p.c, fmaxf(ai - an[c4].z, 0.f));
                    sv[4 * c4 + 3] = fmaf(__uint_as_float(rn[4 * c4 + 3]), p.c, fmaxf(ai - an[c4].w, 0.f));
                }
            } else {
#pragma unroll
                for (int c = 0; c < EL; ++c) sv[c] = __uint_as_float(rn[c]) * p.c;
                if (p.bias_mode == 2 && row_ok) {
#pragma unroll
                    for (int c = 0; c < EL; ++c) {
                        const int key = j * FA_BKV + part * EL + c;
                        if (key < N) sv[c] += __ldg(dense_row + (size_t)key * p.H) * p.c;
                    }
                }
            }
            if ((j + 1) * FA_BKV > N) {                // ragged last tile: keys >= N do not exist
#pragma unroll
                for (int c = 0; c < EL; ++c)
                    if (j * FA_BKV + part * EL + c >= N) sv[c] = -INFINITY;
            }
            float mx0 = fmaxf(sv[0], sv[1]), mx1 = fmaxf(sv[2], sv[3]), mx2 = fmaxf(sv[4], sv[5]), mx3 = fmaxf(sv[6], sv[7]);
#pragma unroll
            for (int c = 8; c < EL; c += 4) {
                mx0 = fmaxf(mx0, sv[c]); mx1 = fmaxf(mx1, sv[c + 1]); mx2 = fmaxf(mx2, sv[c + 2]); mx3 = fmaxf(mx3, sv[c + 3]);
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));     // -inf when this slice has no valid key
            float* xr = xchg + ((j & 1) * FA_BQ + row) * NPART;
            VOG_PROF(2)
            xr[part] = mx;
            row_sync();
            VOG_PROF(3)
            float m_new = m_run;
#pragma unroll
            for (int q = 0; q < NPART; ++q) m_new = fmaxf(m_new, xr[q]);
            if (j == 0) {
                m_run = m_new;
            } else {
                const bool grow = (m_new - m_run) > FA_RESCALE_T;
                if (__any_sync(0xffffffffu, grow)) {
                    // rescale this thread's slice of the accumulator row; PV_{j-1} must have landed first
                    const float alpha = grow ? fast_exp2(m_run - m_new) : 1.f;
                    if (grow) { m_run = m_new; l_run *= alpha; }
                    mbar_wait(pv_done((j - 1) & 1), (uint32_t)(((j - 1) >> 1) & 1));
                    tc_fence_after();
                    for (int c0 = part * ocols; c0 < (part + 1) * ocols; c0 += 16) {
                        uint32_t o[16];
                        tmem_ld16(tmem_o + lane_addr + c0, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
                        tmem_st16(tmem_o + lane_addr + c0, o);
                    }
                    tmem_wait_st();
                }
            }
            if (j + 1 < T && p.bias_mode == 1) load_bias(j + 1);   // sv holds tile j now: an is free again
            VOG_PROF(7)
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < EL; c += 4) {
                sv[c] = fast_exp2(sv[c] - m_run); s0 += sv[c];
                sv[c + 1] = fast_exp2(sv[c + 1] - m_run); s1 += sv[c + 1];
                sv[c + 2] = fast_exp2(sv[c + 2] - m_run); s2 += sv[c + 2];
                sv[c + 3] = fast_exp2(sv[c + 3] - m_run); s3 += sv[c + 3];
            }
            l_run += (s0 + s1) + (s2 + s3);
            VOG_PROF(4)
            if (p.drop_p > 0.f) {
                // dropout AFTER the softmax: the row sum keeps every probability, the PV product only the kept ones,
                // scaled by 1/(1-p); the mask is regenerated by the backward from the same counters
                const float inv_keep = drop_inv_keep8(p.drop_p);
                const uint32_t thr = drop_threshold8(p.drop_p);
#pragma unroll
                for (int c = 0; c < EL; c += 16) {
                    uint32_t rnd[4];
                    attn_rand8x16(p.seed, (uint32_t)bh, (uint32_t)qi, (uint32_t)(j * FA_BKV + part * EL + c), rnd);
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        sv[c + e] = ((rnd[e >> 2] >> (8 * (e & 3))) & 0xffu) >= thr ? sv[c + e] * inv_keep : 0.f;
                }
            }
            VOG_PROF(5)
            // P_j (bf16 pairs) overwrites S_j in place: this thread's EL probabilities -> EL/2 words at columns
            // part*EL/2 of buffer j&1.  Every S_j column was read (all threads of the row passed row_sync)
            {
                uint32_t pw[EL / 2];
#pragma unroll
                for (int i = 0; i < EL / 2; ++i) pw[i] = pack_bf16(sv[2 * i], sv[2 * i + 1]);
                if constexpr (EL == 32) tmem_st16(tmem_s0 + lane_addr + sb * FA_BKV + part * (EL / 2), pw);
                else tmem_st8(tmem_s0 + lane_addr + sb * FA_BKV + part * (EL / 2), pw);
                tmem_wait_st();
            }
            tc_fence_before();
            mbar_arrive(p_full(sb));
            VOG_PROF(6)
            // S_{j+1} (other buffer) was issued right after PV_{j-1}: complete by now in steady state
            if (j + 1 < T) load_scores(j + 1);
        }
#ifdef VOG_ATTN_PROFILE
        if (do_prof) { for (int i = 0; i < 8; ++i) p.prof[i] = pc[i]; p.prof[12] = T; }
#endif
        // ---- epilogue: O / l  (row sum = the partial sums of the row's threads)
        {
            float* xr = xchg + ((T & 1) * FA_BQ + row) * NPART;
            xr[part] = l_run;
            row_sync();
            l_run = 0.f;
#pragma unroll
            for (int q = 0; q < NPART; ++q) l_run += xr[q];
        }
        mbar_wait(pv_done((T - 1) & 1), (uint32_t)(((T - 1) >> 1) & 1));
        tc_fence_after();
        const float inv_l = 1.f / l_run;
        if (p.lse != nullptr && part == 0 && row_ok) p.lse[(size_t)bh * N + qi] = m_run + log2f(l_run);
        const int n_pv = (dh + 15) & ~15;
        for (int c0 = part * ocols; c0 < (part + 1) * ocols; c0 += 16) {
            uint32_t o[16];
            float v[16];
            if (c0 < n_pv) {
                tmem_ld16(tmem_o + lane_addr + c0, o);
                tmem_wait_ld();
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = (c0 + c < n_pv) ? __uint_as_float(o[c]) * inv_l : 0.f;
            if (row_ok) {
                const size_t off = ((size_t)bt * N + qi) * p.ldo + (size_t)h * dhp + c0;
                if (p.out_kind == 1) {
                    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
#pragma unroll
                    for (int c = 0; c < 16; c += 8)
                        *reinterpret_cast<uint4*>(dst + c) =
                            make_uint4(pack_bf16(v[c], v[c + 1]), pack_bf16(v[c + 2], v[c + 3]),
                                       pack_bf16(v[c + 4], v[c + 5]), pack_bf16(v[c + 6], v[c + 7]));
                } else {
                    float* dst = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
                    for (int c = 0; c < 16; c += 4)
                        *reinterpret_cast<float4*>(dst + c) =
                            make_float4(to_tf32(v[c]), to_tf32(v[c + 1]), to_tf32(v[c + 2]), to_tf32(v[c + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CW + 1) tmem_dealloc(tmem_base, p.tmem_cols);
    if (C > 1) cluster_sync_all();                     // no CTA leaves while a peer may still signal its barriers
}


// ak_seq[bt*H + h, key] = c * a[(bt*nbox + key % nbox), h] for key < N, 0 up to the 64-padded row end:
// the per-key factor of the rank-1 bias in the order and scaling the softmax warps consume it
__global__ void bias_expand_kernel(const float* __restrict__ a, float* __restrict__ ak, int Bt, int N, int H,
                                   int nbox, int ld, float c)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= (long long)Bt * H * ld) return;
    const int key = (int)(idx % ld);
    const int bh = (int)(idx / ld);
    const int bt = bh / H, h = bh % H;
    ak[idx] = key < N ? a[((size_t)bt * nbox + key % nbox) * H + h] * c : 0.f;
}

long long tc_attn_workspace_bytes(int Bt, int N, int H)
{
    return (long long)Bt * H * round_up(N, FA_BKV) * 4;
}
int tc_attn_key_ld(int N) { return round_up(N, FA_BKV); }
float tc_attn_key_scale(float inv_scale) { return inv_scale * 1.4426950408889634f; }

static thread_local long long* g_attn_prof = nullptr;
static thread_local int g_attn_impl = 2;             // 1 = Q/P through smem (v1); Q/P in tensor memory with 2 (v2) / 4 (v3) softmax threads per row
static thread_local int g_attn_cluster = 0;          // 0 = default (no cluster); 2 / 4 = multicast K/V loads across query tiles
void tc_attn_set_impl(int impl) { g_attn_impl = (impl >= 1 && impl <= 3) ? impl : 2; }
void tc_attn_set_cluster(int c) { g_attn_cluster = c; }
void tc_attn_set_prof(long long* buf) { g_attn_prof = buf; }

int tc_attn(const void* q, const void* k, const void* v, int Bt, int N, int H, int dhp,
            const int* dh, float inv_scale, int bias_mode, const float* a, int nbox, const float* bpe,
            const float* dense, void* out, long long ldo, int out_kind, void* workspace,
            long long workspace_bytes, cudaStream_t st, float* lse, float drop_p, unsigned long long seed)
{
    if (Bt == 0 || N == 0) return 0;
    VOG_REQUIRE((lse == nullptr && drop_p == 0.f) || g_attn_impl >= 2, "tc_attn: lse / dropout need the v2 kernel");
    VOG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "tc_attn: dropout probability %f", (double)drop_p);
    VOG_REQUIRE(H >= 1 && H <= VOG_MAX_HEADS, "tc_attn: H=%d out of range", H);
    VOG_REQUIRE(dhp == 64 || dhp == 128 || dhp == 192 || dhp == 256, "tc_attn: dhp=%d must be 64/128/192/256", dhp);
    VOG_REQUIRE(Bt <= 65535 && H <= 65535, "tc_attn: grid too large");
    VOG_REQUIRE(out_kind == 1 || out_kind == 2, "tc_attn: bad out_kind");
    VOG_REQUIRE(ldo >= (long long)H * dhp && (ldo * (out_kind == 1 ? 2 : 4)) % 16 == 0, "tc_attn: bad ldo");
    VOG_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "tc_attn: output must be 16-byte aligned");
    VOG_REQUIRE((bias_mode != 1 && bias_mode != 3) || (a && bpe && nbox > 0), "tc_attn: rank-1 bias needs a, bpe, nbox");
    VOG_REQUIRE(bias_mode != 3 || (lse == nullptr && drop_p == 0.f), "tc_attn: pre-expanded key factors are an inference-path option");
    VOG_REQUIRE(bias_mode != 2 || dense, "tc_attn: dense bias pointer missing");
    AttnParams p;
    p.Bt = Bt; p.N = N; p.H = H; p.dhp = dhp;
    for (int h = 0; h < H; ++h) {
        VOG_REQUIRE(dh[h] >= 1 && dh[h] <= dhp, "tc_attn: head dim %d does not fit dhp=%d", dh[h], dhp);
        p.dh[h] = dh[h];
    }
    p.c = tc_attn_key_scale(inv_scale);
    // bias_mode 3 = rank-1 whose key factors the caller already expanded into the workspace (vog_pe_project_expand)
    const bool pre_expanded = bias_mode == 3;
    if (pre_expanded) bias_mode = 1;
    p.bias_mode = bias_mode; p.a = a; p.nbox = nbox > 0 ? nbox : 1; p.bpe = bpe; p.dense = dense;
    p.out = out; p.ldo = ldo; p.out_kind = out_kind;
    p.prof = g_attn_prof;
    p.lse = lse; p.drop_p = drop_p; p.seed = seed;
    p.ak_seq = nullptr; p.ak_ld = round_up(N, FA_BKV);
    if (bias_mode == 1) {
        VOG_REQUIRE(workspace && workspace_bytes >= tc_attn_workspace_bytes(Bt, N, H) &&
                    (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                    "tc_attn: rank-1 bias needs a 16-byte aligned workspace of tc_attn_workspace_bytes()");
        p.ak_seq = reinterpret_cast<const float*>(workspace);
        if (!pre_expanded) {
            const long long n = (long long)Bt * H * p.ak_ld;
            VOG_CUDA(launch_pdl(bias_expand_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, a,
                                reinterpret_cast<float*>(workspace), Bt, N, H, p.nbox, p.ak_ld, p.c));
            if (check_launch("bias_expand")) return -1;
        }
    }
    p.q = reinterpret_cast<const __nv_bfloat16*>(q);
    if (g_attn_impl >= 2) {
        const int npart = g_attn_impl == 2 ? 2 : 4;
        p.tmem_cols = 512;
        const int qtiles = cdiv(N, FA_BQ);
        // measured on B200 (profiles/r1/attention.md): multicast clusters LOSE 13-30 % here - the limiter is each SM's
        // own shared-memory fill rate (every CTA still receives the whole tile), not L2 bandwidth, and the CTAs of a
        // cluster then run in lock step.  The path stays available for experiments (vog_debug_attn_cluster).
        const int C = g_attn_cluster > 0 ? g_attn_cluster : 1;
        VOG_REQUIRE((C == 1 || C == 2 || C == 4) && qtiles % C == 0, "tc_attn: cluster size %d does not divide %d query tiles", C, qtiles);
        p.cluster = C;
        CUtensorMap tk2, tv2;
        const uint64_t BH2 = (uint64_t)Bt * H;
        uint64_t dq2[3] = {(uint64_t)dhp, (uint64_t)N, BH2};
        uint64_t sq2[2] = {(uint64_t)dhp * 2, (uint64_t)N * dhp * 2};
        uint32_t bk2[3] = {64, (uint32_t)(FA_BKV / C), 1};       // every CTA of a cluster loads a 64/C-key slice
        if (make_tmap(&tk2, k, 2, 1, 3, dq2, sq2, bk2)) return -1;
        if (make_tmap(&tv2, v, 2, 1, 3, dq2, sq2, bk2)) return -1;
        const int fixed2 = 384 /*barriers*/ + 2 * FA_BQ * npart * 4 /*row-max exchange*/;
        const int stage_bytes2 = 2 * FA_BKV * dhp * 2;
        int stages2 = (227 * 1024 - fixed2) / stage_bytes2;
        if (stages2 > FA_MAX_STAGES) stages2 = FA_MAX_STAGES;
        p.stages = stages2;
        const size_t smem2 = (size_t)fixed2 + (size_t)stages2 * stage_bytes2;
        if (npart == 2) VOG_CUDA(cudaFuncSetAttribute(tc_attn2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        else VOG_CUDA(cudaFuncSetAttribute(tc_attn2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(qtiles, H, Bt);
        cfg.blockDim = dim3(128 * npart + 128, 1, 1);
        cfg.dynamicSmemBytes = smem2;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
        cfg.attrs = attr; cfg.numAttrs = 2;
        if (npart == 2) VOG_CUDA(cudaLaunchKernelEx(&cfg, tc_attn2_kernel<2>, tk2, tv2, p));
        else VOG_CUDA(cudaLaunchKernelEx(&cfg, tc_attn2_kernel<4>, tk2, tv2, p));
        return check_launch("tc_attn2");
    }
    const int cols = dhp + 3 * FA_BKV;
    p.tmem_cols = cols <= 256 ? 256 : 512;

    CUtensorMap tq, tk, tv;
    const uint64_t BH = (uint64_t)Bt * H;
    uint64_t dq[3] = {(uint64_t)dhp, (uint64_t)N, BH};
    uint64_t sq[2] = {(uint64_t)dhp * 2, (uint64_t)N * dhp * 2};
    uint32_t bq[3] = {64, FA_BQ, 1}, bk[3] = {64, FA_BKV, 1};
    if (make_tmap(&tq, q, 2, 1, 3, dq, sq, bq)) return -1;
    if (make_tmap(&tk, k, 2, 1, 3, dq, sq, bk)) return -1;
    if (make_tmap(&tv, v, 2, 1, 3, dq, sq, bk)) return -1;

    const int fixed = FA_BQ * dhp * 2 + 2 * FA_BQ * 128 + 384 /*barriers*/ + 2 * FA_BQ * 2 * 4 /*pair exchange*/;
    const int stage_bytes = 2 * FA_BKV * dhp * 2;                  // one K slot + one V^T slot
    int stages = (227 * 1024 - fixed) / stage_bytes;
    if (stages > FA_MAX_STAGES) stages = FA_MAX_STAGES;
    VOG_REQUIRE(stages >= 2, "tc_attn: not enough shared memory for dhp=%d", dhp);
    p.stages = stages;
    const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
    VOG_CUDA(cudaFuncSetAttribute(tc_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cdiv(N, FA_BQ), H, Bt);
    tc_attn_kernel<<<grid, FA_THREADS, smem, st>>>(tq, tk, tv, p);
    return check_launch("tc_attn");
}

}  // namespace vog
