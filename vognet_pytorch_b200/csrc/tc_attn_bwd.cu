// Attention backward on the tcgen05 tensor cores (training step, compute mode 'bf16').
//
// Gradient of  O = softmax((Q K^T + bias) / sqrt(d_model)) V  with the rank-1 relative-position bias
// bias_ij = relu(a_i - a_j + b)  (code/transformer_code.py:41-50,136-160; code/mdl_vog.py:477-488;
// utils/mdl_srl_utils.py:30-69), as torch autograd derives it for the reference from utils/trn_utils.py:504
// (loss.backward()).  The forward kernel keeps only the row log-sum-exp; the backward recomputes the
// probabilities on the tensor cores:
//
//   delta_i   = sum_c dO_ic O_ic                                          attn_delta_kernel
//   K_A       S = Q K^T and dP = dO V^T as TWO accumulators of one tcgen05 pipeline per 128 x 128 tile;
//             epilogue P = exp2(c S + bias - lse), dS = P o (dP - delta) / sqrt(d_model); P and dS are written
//             once as bf16 [Bt*H, Npad, Npad]; the bias gradient is reduced on the fly: row sums
//             (d a_i), column sums (-d a_j) over the entries where the relu is open      tc_attn_bwd_sdp_kernel
//   K_B       dQ = dS K          (A K-major, B = K in its natural [key][dh] layout = MN-major)   tc_bgemm_kernel<false>
//   K_C       dK = dS^T Q        (A = dS MN-major: the contraction index is its slow index)      tc_bgemm_kernel<true>
//   K_D       dV = P^T dO                                                                      tc_bgemm_kernel<true>
//   fold      d a[box], d b_pe  from the per-token row / column sums                          attn_dbias_fold_kernel
//
// Nothing is transposed in HBM; dQ | dK | dV land as bf16 in one [Bt*N, 3*H*dhp] matrix whose column order is the
// row order of the packed Wq|Wk|Wv operand, so the projection's weight / input gradients are two plain GEMMs.
// Dropout on the probabilities (code/transformer_code.py:153) is regenerated from the same counter-based stream
// as the forward (attn_keep()).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "philox.cuh"

namespace vog {

using namespace tc;

constexpr int AB_T = 128;                          // tile edge: 128 queries x 128 keys
constexpr int AB_EPI_WARPS = 8;                    // two per TMEM lane quarter: 64 key columns each, 16 at a time
constexpr int AB_WCOLS = AB_T / (AB_EPI_WARPS / 4);  // key columns per epilogue warp
constexpr int AB_THREADS = 64 + 32 * AB_EPI_WARPS;
constexpr int AB_STAGES = 3;
constexpr int AB_CHUNK = AB_T * 128;               // [128 rows x 64 bf16], 128B-swizzled
constexpr int AB_RES = 4 * AB_CHUNK;               // resident Q (and dO) tile: up to dhp = 256

struct AttnBwdParams {
    int Bt, N, H, dhp, Npad;
    int qtiles, ktiles, nkb;
    float c, inv_scale;                            // log2(e)/sqrt(d_model), 1/sqrt(d_model)
    int bias_mode;                                 // 0 none, 1 rank-1
    const float* a; int nbox; const float* bpe;    // [Bt*nbox, H], [H]
    const float* ak; int ak_ld;                    // [Bt*H, ak_ld]: c * a[key % nbox], zero beyond N (ak_ld >= ktiles*128)
    const float* lse; const float* delta;          // [Bt*H, N] (lse in the log2 domain of the forward)
    __nv_bfloat16* P; __nv_bfloat16* dS;           // [Bt*H, Npad, Npad]
    float* drow; float* dcol;                      // [Bt*H, N], accumulated
    float drop_p; unsigned long long seed;
    uint32_t idesc;
};

__device__ __forceinline__ float ab_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ab_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * AB_EPI_WARPS) : "memory"); }

__global__ void __launch_bounds__(AB_THREADS, 1)
tc_attn_bwd_sdp_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                       const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do,
                       const AttnBwdParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) { if (threadIdx.x == 0) printf("vog: dynamic smem not 1024-aligned\n"); __trap(); }
    const uint32_t q_res = smem_base, do_res = smem_base + AB_RES;
    const uint32_t ring = smem_base + 2 * AB_RES;
    constexpr uint32_t stage_bytes = 2 * AB_CHUNK;
    const uint32_t bar_off = 2 * AB_RES + AB_STAGES * stage_bytes;
    const uint32_t bar_base = smem_base + bar_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (AB_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * AB_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * AB_STAGES + 2 + a); };
    const uint32_t qdo_full = bar_base + 8u * (2 * AB_STAGES + 4);
    const uint32_t qdo_empty = bar_base + 8u * (2 * AB_STAGES + 5);
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + bar_off + 8 * (2 * AB_STAGES + 6));
    float* colsc = reinterpret_cast<float*>(smem_raw + bar_off + 128);           // [4 lane quarters][128 columns]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do);
        for (int s = 0; s < AB_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), AB_EPI_WARPS); }
        mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int nitems = p.Bt * p.H * p.qtiles;
    const int nkb = p.nkb;

    if (warp == 0) {
        // ================= TMA producer =================
        int s = 0; uint32_t ph = 0; int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            const int bh = item / p.qtiles, qt = item - bh * p.qtiles;
            const int bt = bh / p.H, h = bh - bt * p.H;
            mbar_wait(qdo_empty, (uint32_t)((it & 1) ^ 1));        // the previous item's MMAs have read Q / dO
            if (elect_one()) {
                mbar_arrive_expect_tx(qdo_full, 2u * nkb * AB_CHUNK);
                for (int kb = 0; kb < nkb; ++kb) {
                    tma_load_3d(q_res + kb * AB_CHUNK, &tma_q, qdo_full, kb * 64, qt * AB_T, bh);
                    tma_load_3d(do_res + kb * AB_CHUNK, &tma_do, qdo_full, h * p.dhp + kb * 64, qt * AB_T, bt);
                }
            }
            __syncwarp();
            for (int kt = 0; kt < p.ktiles; ++kt) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                        tma_load_3d(ring + s * stage_bytes, &tma_k, full_bar(s), kb * 64, kt * AB_T, bh);
                        tma_load_3d(ring + s * stage_bytes + AB_CHUNK, &tma_v, full_bar(s), kb * 64, kt * AB_T, bh);
                    }
                    __syncwarp();
                    if (++s == AB_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: S and dP accumulate side by side =================
        int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0; int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            mbar_wait(qdo_full, (uint32_t)(it & 1));
            tc_fence_after();
            for (int kt = 0; kt < p.ktiles; ++kt) {
                mbar_wait(tempty_bar(acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t t_s = tmem_base + acc * 256, t_dp = t_s + 128;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint32_t qa = umma_desc_lo(q_res + kb * AB_CHUNK), da = umma_desc_lo(do_res + kb * AB_CHUNK);
                    const uint32_t kd = umma_desc_lo(ring + s * stage_bytes), vd = umma_desc_lo(ring + s * stage_bytes + AB_CHUNK);
                    if (elect_one()) {
                        if (kb == 0) {
                            umma_bf16_lo<false>(t_s, qa, kd, p.idesc);
                            umma_bf16_lo<false>(t_dp, da, vd, p.idesc);
                        } else {
                            umma_bf16_lo<true>(t_s, qa, kd, p.idesc);
                            umma_bf16_lo<true>(t_dp, da, vd, p.idesc);
                        }
#pragma unroll
                        for (int k = 1; k < 4; ++k) {
                            umma_bf16_lo<true>(t_s, qa + 2 * k, kd + 2 * k, p.idesc);
                            umma_bf16_lo<true>(t_dp, da + 2 * k, vd + 2 * k, p.idesc);
                        }
                        umma_commit(empty_bar(s));
                    }
                    __syncwarp();
                    if (++s == AB_STAGES) { s = 0; ph ^= 1; }
                }
                if (elect_one()) {
                    umma_commit(tfull_bar(acc));
                    if (kt == p.ktiles - 1) umma_commit(qdo_empty);
                }
                __syncwarp();
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
            }
        }
    } else {
        // ================= epilogue: thread = query row; warp (g, cq) = TMEM lane quarter g, key columns
        // [AB_WCOLS*cq, +AB_WCOLS), processed 16 at a time (one Philox call, one 32-byte store per tensor; no spills:
        // local-memory reloads miss the L1 that the streaming stores keep flushing) ====
        const int g = warp & 3, cq = (warp - 2) >> 2;
        const int et = threadIdx.x - 64;                         // 0..511
        const uint32_t lane_addr = (uint32_t)(32 * g) << 16;
        const bool rel = p.bias_mode == 1;
        const bool drop = p.drop_p > 0.f;
        const float inv_keep = drop ? drop_inv_keep8(p.drop_p) : 1.f;
        const uint32_t thr = drop_threshold8(p.drop_p);
        const float k2 = p.inv_scale * inv_keep;
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int bh = item / p.qtiles, qt = item - bh * p.qtiles;
            const int bt = bh / p.H, h = bh - bt * p.H;
            const int qi = qt * AB_T + 32 * g + lane;
            const bool row_ok = qi < p.N;
            float lse = 0.f, dlt = 0.f, ai = 0.f;
            if (row_ok) {
                lse = __ldg(p.lse + (size_t)bh * p.N + qi);
                dlt = __ldg(p.delta + (size_t)bh * p.N + qi) * p.inv_scale;
                if (rel) ai = (__ldg(p.a + ((size_t)bt * p.nbox + qi % p.nbox) * p.H + h) + __ldg(p.bpe + h)) * p.c;
            }
            float rowsum0 = 0.f, rowsum1 = 0.f;
            const size_t rowoff = ((size_t)bh * p.Npad + qi) * p.Npad;
            const float* akrow = p.ak + (size_t)bh * p.ak_ld + AB_WCOLS * cq;
            float4 an[4];                                        // bias factors of the NEXT half: loaded one half ahead
            if (rel) {
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) an[j4] = __ldg(reinterpret_cast<const float4*>(akrow) + j4);
            }
            for (int kt = 0; kt < p.ktiles; ++kt) {
                mbar_wait(tfull_bar(acc), acc_ph);
                tc_fence_after();
                const uint32_t t_s = tmem_base + lane_addr + acc * 256 + AB_WCOLS * cq, t_dp = t_s + 128;
                const bool edge = (qt == p.qtiles - 1 || kt == p.ktiles - 1) && p.N != p.Npad;
#pragma unroll 1
                for (int hh = 0; hh < AB_WCOLS / 16; ++hh) {
                    const int key0 = kt * AB_T + AB_WCOLS * cq + 16 * hh;
                    uint32_t rs[16], rp[16];
                    tmem_ld16(t_s + 16 * hh, rs);
                    tmem_ld16(t_dp + 16 * hh, rp);
                    float aj[16];
                    if (rel) {
#pragma unroll
                        for (int j4 = 0; j4 < 4; ++j4) { aj[4 * j4] = an[j4].x; aj[4 * j4 + 1] = an[j4].y; aj[4 * j4 + 2] = an[j4].z; aj[4 * j4 + 3] = an[j4].w; }
                        // next half (next tile's first half after the second one); the row is padded to Npad
                        const int nk = hh + 1 < AB_WCOLS / 16 ? key0 + 16 : (kt + 1 < p.ktiles ? key0 + AB_T - AB_WCOLS + 16 : -1);
                        if (nk >= 0) {
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4)
                                an[j4] = __ldg(reinterpret_cast<const float4*>(p.ak + (size_t)bh * p.ak_ld + nk) + j4);
                        }
                    }
                    uint32_t rnd[4];
                    if (drop) attn_rand8x16(p.seed, (uint32_t)bh, (uint32_t)qi, (uint32_t)key0, rnd);
                    tmem_wait_ld();
                    float gv[16];
                    uint32_t pw[8], gw[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        float pj[2], gj[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int jj = j + e;
                            const float bdiff = rel ? ai - aj[jj] : 0.f;
                            const float t = fmaf(__uint_as_float(rs[jj]), p.c, fmaxf(bdiff, 0.f)) - lse;
                            float pr = ab_exp2(t);
                            if (edge) pr = (row_ok && key0 + jj < p.N) ? pr : 0.f;
                            float base = fmaf(__uint_as_float(rp[jj]), k2, -dlt);   // (dP o mask / keep - delta) / sqrt(d_model)
                            float pstore = pr;
                            if (drop) {
                                const bool keep = ((rnd[jj >> 2] >> (8 * (jj & 3))) & 0xffu) >= thr;
                                pstore = keep ? pr * inv_keep : 0.f;
                                base = keep ? base : -dlt;
                            }
                            const float gg = pr * base;
                            pj[e] = pstore; gj[e] = gg;
                            const float gm = bdiff > 0.f ? gg : 0.f;
                            gv[jj] = gm;
                            if (e == 0) rowsum0 += gm; else rowsum1 += gm;
                        }
                        pw[j >> 1] = pack_bf16(pj[0], pj[1]);
                        gw[j >> 1] = pack_bf16(gj[0], gj[1]);
                    }
                    st_global_v8(p.P + rowoff + key0, pw);            // 16 bf16 = one 32-byte sector per thread
                    st_global_v8(p.dS + rowoff + key0, gw);
                    if (rel) {
                        // column sums over the warp's 32 rows: butterfly reduce-scatter of 16 values, lanes 2c and 2c+1
                        // end with column key0 + c
#pragma unroll
                        for (int off = 16; off >= 2; off >>= 1) {
#pragma unroll
                            for (int k = 0; k < off / 2; ++k) {
                                const bool up = (lane & off) != 0;
                                const float send = up ? gv[k] : gv[k + off / 2];
                                const float keepv = up ? gv[k + off / 2] : gv[k];
                                gv[k] = keepv + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                        gv[0] += __shfl_xor_sync(0xffffffffu, gv[0], 1);
                        if ((lane & 1) == 0) colsc[g * AB_T + AB_WCOLS * cq + 16 * hh + (lane >> 1)] = gv[0];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(acc));           // accumulators are free for the tile after next
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
                if (rel) {
                    ab_epi_bar();
                    if (et < AB_T) {
                        const int key = kt * AB_T + et;
                        const float v = (colsc[et] + colsc[AB_T + et]) + (colsc[2 * AB_T + et] + colsc[3 * AB_T + et]);
                        if (key < p.N) atomicAdd(p.dcol + (size_t)bh * p.N + key, v);
                    }
                    ab_epi_bar();
                }
            }
            if (rel && row_ok) atomicAdd(p.drow + (size_t)bh * p.N + qi, rowsum0 + rowsum1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- batched GEMM of the backward: C[bh] (M x dhp) = op(A[bh]) . B[bh], B MN-major ------------------------------
constexpr int BG_BM = 128;
constexpr int BG_BK = 64;
constexpr int BG_SUB = 64 * 128;                   // [64 rows x 128 B] sub-tile
constexpr int BG_EPI_WARPS = 8;
constexpr int BG_THREADS = 64 + 32 * BG_EPI_WARPS;
constexpr int BG_MAX_STAGES = 6;

struct BGemmParams {
    int BH, H, N, dhp, mtiles, kblocks, nsub, stages;
    int b_hcol;                                    // B inner-dimension offset per head (dO: dhp, Q / K: 0)
    int b_zbt;                                     // B outer coordinate: 1 = bt, 0 = bt*H + h
    __nv_bfloat16* out; long long ldo; int out_col0;        // out[(bt*N + m)*ldo + out_col0 + h*dhp + c]
    uint32_t idesc, tmem_cols;
};

template <bool kAMN>
__global__ void __launch_bounds__(BG_THREADS, 1)
tc_bgemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const BGemmParams p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if ((smem_base & 1023u) != 0) { if (threadIdx.x == 0) printf("vog: dynamic smem not 1024-aligned\n"); __trap(); }
    const uint32_t a_bytes = 2 * BG_SUB, b_bytes = p.nsub * BG_SUB;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t bar_off = p.stages * stage_bytes;
    const uint32_t bar_base = smem_base + bar_off;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (BG_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * BG_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * BG_MAX_STAGES + 2 + a); };
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + bar_off + 8 * (2 * BG_MAX_STAGES + 4));
    auto a_smem = [&](int s) { return smem_base + s * stage_bytes; };
    auto b_smem = [&](int s) { return smem_base + s * stage_bytes + a_bytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), BG_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int nitems = p.BH * p.mtiles;

    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int bh = item / p.mtiles, mt = item - bh * p.mtiles;
            const int bt = bh / p.H, h = bh - bt * p.H;
            const int bz = p.b_zbt ? bt : bh, bc = p.b_hcol * h;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                    if constexpr (kAMN) {
                        // A stored [contraction row][m]: two {64 m x 64 rows} boxes
                        tma_load_3d(a_smem(s), &tma_a, full_bar(s), mt * BG_BM, kb * BG_BK, bh);
                        tma_load_3d(a_smem(s) + BG_SUB, &tma_a, full_bar(s), mt * BG_BM + 64, kb * BG_BK, bh);
                    } else {
                        // A stored [m][contraction]: one {64 contraction x 128 m} box
                        tma_load_3d(a_smem(s), &tma_a, full_bar(s), kb * BG_BK, mt * BG_BM, bh);
                    }
                    for (int i = 0; i < p.nsub; ++i)
                        tma_load_3d(b_smem(s) + i * BG_SUB, &tma_b, full_bar(s), bc + i * 64, kb * BG_BK, bz);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0;
        const uint32_t lbo = ((uint32_t)BG_SUB >> 4) << 16;          // MN-major: next 64 MN elements one sub-tile further
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * p.dhp;
            for (int kb = 0; kb < p.kblocks; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a_lo = kAMN ? ((((a_smem(s)) >> 4) & 0x3FFF) | lbo) : umma_desc_lo(a_smem(s));
                const uint32_t b_lo = (((b_smem(s)) >> 4) & 0x3FFF) | lbo;
                constexpr uint32_t a_step = kAMN ? (2048 >> 4) : 2;   // K = 16: 16 rows x 128 B (MN-major) / 32 B (K-major)
                if (elect_one()) {
                    if (kb == 0) umma_bf16_lo<false>(d_tmem, a_lo, b_lo, p.idesc);
                    else umma_bf16_lo<true>(d_tmem, a_lo, b_lo, p.idesc);
#pragma unroll
                    for (int k = 1; k < BG_BK / 16; ++k)
                        umma_bf16_lo<true>(d_tmem, a_lo + k * a_step, b_lo + k * (2048 >> 4), p.idesc);
                    umma_commit(empty_bar(s));
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) umma_commit(tfull_bar(acc));
            __syncwarp();
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1;
        }
    } else {
        const int g = warp & 3, half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int bh = item / p.mtiles, mt = item - bh * p.mtiles;
            const int bt = bh / p.H, h = bh - bt * p.H;
            const int m = mt * BG_BM + 32 * g + lane;
            const uint32_t t_acc = tmem_base + ((uint32_t)(32 * g) << 16) + acc * p.dhp;
            __nv_bfloat16* orow = p.out + ((size_t)bt * p.N + m) * p.ldo + p.out_col0 + (size_t)h * p.dhp;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            uint32_t r[32];
#pragma unroll 1
            for (int c0 = 32 * half; c0 < p.dhp; c0 += 64) {
                tmem_ld32(t_acc + c0, r);
                tmem_wait_ld();
                if (m < p.N) {
                    uint32_t w[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) w[i] = pack_bf16(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                    st_global_v8(orow + c0, w);
                    st_global_v8(orow + c0 + 16, w + 8);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// delta[bh, i] = sum_c dO[bt*N + i, h*dhp + c] * O[bt*N + i, h*dhp + c]      (one warp per (bt, i, h))
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ dout, long long lddo,
                  float* __restrict__ delta, int Bt, int N, int H, int dhp)
{
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)Bt * N * H) return;
    const int h = (int)(w % H);
    const long long row = w / H;                    // bt*N + i
    const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(o + row * ldo + (size_t)h * dhp);
    const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(dout + row * lddo + (size_t)h * dhp);
    float s = 0.f;
    for (int c = lane; c < dhp / 2; c += 32) {
        const float2 x = __bfloat1622float2(a[c]), y = __bfloat1622float2(b[c]);
        s = fmaf(x.x, y.x, s);
        s = fmaf(x.y, y.y, s);
    }
    s = warp_sum(s);
    if (lane == 0) {
        const int bt = (int)(row / N), i = (int)(row % N);
        delta[((size_t)bt * H + h) * N + i] = s;
    }
}

// per-key factor of the rank-1 bias, as the forward's bias_expand_kernel but with the backward's row length
__global__ void attn_bwd_bias_expand_kernel(const float* __restrict__ a, float* __restrict__ ak, int Bt, int N, int H,
                                            int nbox, int ld, float c)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)Bt * H * ld) return;
    const int key = (int)(idx % ld);
    const int bh = (int)(idx / ld);
    const int bt = bh / H, h = bh % H;
    ak[idx] = key < N ? a[((size_t)bt * nbox + key % nbox) * H + h] * c : 0.f;
}

// d a[bt*nbox + i % nbox, h] += drow[bh, i] - dcol[bh, i];  d b_pe[h] += sum drow
__global__ void __launch_bounds__(256)
attn_dbias_fold_kernel(const float* __restrict__ drow, const float* __restrict__ dcol, float* __restrict__ da,
                       float* __restrict__ dbpe, int Bt, int N, int H, int nbox)
{
    __shared__ float hb[VOG_MAX_HEADS];
    if (threadIdx.x < VOG_MAX_HEADS) hb[threadIdx.x] = 0.f;
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)Bt * H * N) {
        const int i = (int)(idx % N);
        const int bh = (int)(idx / N);
        const int bt = bh / H, h = bh % H;
        const float r = drow[idx], cv = dcol[idx];
        atomicAdd(da + ((size_t)bt * nbox + i % nbox) * H + h, r - cv);
        atomicAdd(&hb[h], r);
    }
    __syncthreads();
    if (threadIdx.x < H && hb[threadIdx.x] != 0.f) atomicAdd(dbpe + threadIdx.x, hb[threadIdx.x]);
}

static inline long long al256(long long x) { return (x + 255) & ~255LL; }

long long tc_attn_bwd_workspace_bytes(int Bt, int N, int H)
{
    const long long BH = (long long)Bt * H, Npad = round_up(N, AB_T);
    return 2 * al256(BH * Npad * Npad * 2) + al256(BH * Npad * 4) + 3 * al256(BH * N * 4);
}

static int launch_bgemm(bool a_mn, const __nv_bfloat16* A, int Npad, const void* B, const uint64_t* bdims,
                        const uint64_t* bstrides, int b_hcol, int b_zbt, int Bt, int N, int H, int dhp, __nv_bfloat16* out,
                        long long ldo, int out_col0, cudaStream_t st)
{
    const int BH = Bt * H;
    CUtensorMap ta, tb;
    uint64_t da[3] = {(uint64_t)Npad, (uint64_t)Npad, (uint64_t)BH};
    uint64_t sa[2] = {(uint64_t)Npad * 2, (uint64_t)Npad * Npad * 2};
    uint32_t boxa_k[3] = {64, 128, 1}, boxa_mn[3] = {64, 64, 1}, boxb[3] = {64, 64, 1};
    if (make_tmap(&ta, A, 2, 1, 3, da, sa, a_mn ? boxa_mn : boxa_k)) return -1;
    if (make_tmap(&tb, B, 2, 1, 3, bdims, bstrides, boxb)) return -1;
    BGemmParams p;
    p.BH = BH; p.H = H; p.N = N; p.dhp = dhp;
    p.mtiles = Npad / BG_BM; p.kblocks = Npad / BG_BK; p.nsub = dhp / 64;
    p.b_hcol = b_hcol; p.b_zbt = b_zbt;
    p.out = out; p.ldo = ldo; p.out_col0 = out_col0;
    p.idesc = umma_idesc(FMT_BF16, BG_BM, dhp) | (1u << 16) | (a_mn ? (1u << 15) : 0u);
    p.tmem_cols = 2 * dhp <= 128 ? 128 : 2 * dhp <= 256 ? 256 : 512;
    const int stage_bytes = (2 + p.nsub) * BG_SUB;
    int stages = (227 * 1024 - 256) / stage_bytes;
    if (stages > BG_MAX_STAGES) stages = BG_MAX_STAGES;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 256;
    const int nitems = BH * p.mtiles;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    const int grid = nitems < sms ? nitems : sms;
    if (a_mn) {
        VOG_CUDA(cudaFuncSetAttribute(tc_bgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_bgemm_kernel<true><<<grid, BG_THREADS, smem, st>>>(ta, tb, p);
    } else {
        VOG_CUDA(cudaFuncSetAttribute(tc_bgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_bgemm_kernel<false><<<grid, BG_THREADS, smem, st>>>(ta, tb, p);
    }
    return check_launch("tc_bgemm");
}

int tc_attn_bwd(const void* q, const void* k, const void* v, const void* o, long long ldo, const void* dout, long long lddo,
                const float* lse, int Bt, int N, int H, int dhp, const int* dh, float inv_scale, int bias_mode,
                const float* a, int nbox, const float* bpe, void* dqkv, long long ldg, float* da, float* dbpe,
                void* workspace, long long workspace_bytes, float drop_p, unsigned long long seed, cudaStream_t st)
{
    if (Bt == 0 || N == 0) return 0;
    VOG_REQUIRE(H >= 1 && H <= VOG_MAX_HEADS, "tc_attn_bwd: H=%d out of range", H);
    VOG_REQUIRE(dhp == 64 || dhp == 128 || dhp == 192 || dhp == 256, "tc_attn_bwd: dhp=%d must be 64/128/192/256", dhp);
    VOG_REQUIRE(bias_mode == 0 || bias_mode == 1, "tc_attn_bwd: bias_mode %d (dense bias: use the fp32x mode)", bias_mode);
    VOG_REQUIRE(bias_mode == 0 || (a && bpe && nbox > 0 && da && dbpe), "tc_attn_bwd: rank-1 bias needs a, bpe, nbox, da, dbpe");
    VOG_REQUIRE(ldo >= (long long)H * dhp && lddo >= (long long)H * dhp && ldg >= 3LL * H * dhp && ldo % 8 == 0 && lddo % 8 == 0 &&
                ldg % 8 == 0, "tc_attn_bwd: bad leading dimension");
    VOG_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "tc_attn_bwd: dropout probability %f", (double)drop_p);
    VOG_REQUIRE(workspace && workspace_bytes >= tc_attn_bwd_workspace_bytes(Bt, N, H) &&
                (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "tc_attn_bwd: workspace too small / not 256-byte aligned");
    for (int h = 0; h < H; ++h) VOG_REQUIRE(dh[h] >= 1 && dh[h] <= dhp, "tc_attn_bwd: head dim %d does not fit dhp=%d", dh[h], dhp);
    const int BH = Bt * H, Npad = round_up(N, AB_T);
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(w); w += al256((long long)BH * Npad * Npad * 2);
    __nv_bfloat16* dS = reinterpret_cast<__nv_bfloat16*>(w); w += al256((long long)BH * Npad * Npad * 2);
    float* ak = reinterpret_cast<float*>(w); w += al256((long long)BH * Npad * 4);
    float* delta = reinterpret_cast<float*>(w); w += al256((long long)BH * N * 4);
    float* drow = reinterpret_cast<float*>(w); w += al256((long long)BH * N * 4);
    float* dcol = reinterpret_cast<float*>(w);

    {   // delta
        const long long warps = (long long)Bt * N * H;
        attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(o), ldo, reinterpret_cast<const __nv_bfloat16*>(dout), lddo, delta, Bt, N, H, dhp);
        if (check_launch("attn_delta")) return -1;
    }
    AttnBwdParams p;
    p.Bt = Bt; p.N = N; p.H = H; p.dhp = dhp; p.Npad = Npad;
    p.qtiles = Npad / AB_T; p.ktiles = Npad / AB_T; p.nkb = dhp / 64;
    p.c = inv_scale * 1.4426950408889634f; p.inv_scale = inv_scale;
    p.bias_mode = bias_mode; p.a = a; p.nbox = nbox > 0 ? nbox : 1; p.bpe = bpe;
    p.ak = ak; p.ak_ld = Npad; p.lse = lse; p.delta = delta; p.P = P; p.dS = dS; p.drow = drow; p.dcol = dcol;
    p.drop_p = drop_p; p.seed = seed;
    p.idesc = umma_idesc(FMT_BF16, AB_T, AB_T);
    if (bias_mode == 1) {
        const long long n = (long long)BH * Npad;
        attn_bwd_bias_expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, ak, Bt, N, H, p.nbox, Npad, p.c);
        if (check_launch("attn_bwd_bias_expand")) return -1;
        VOG_CUDA(cudaMemsetAsync(drow, 0, 2 * al256((long long)BH * N * 4), st));
    }
    CUtensorMap tq, tk, tv, tdo;
    uint64_t dq[3] = {(uint64_t)dhp, (uint64_t)N, (uint64_t)BH};
    uint64_t sq[2] = {(uint64_t)dhp * 2, (uint64_t)N * dhp * 2};
    uint32_t box[3] = {64, AB_T, 1};
    if (make_tmap(&tq, q, 2, 1, 3, dq, sq, box)) return -1;
    if (make_tmap(&tk, k, 2, 1, 3, dq, sq, box)) return -1;
    if (make_tmap(&tv, v, 2, 1, 3, dq, sq, box)) return -1;
    uint64_t ddo[3] = {(uint64_t)H * dhp, (uint64_t)N, (uint64_t)Bt};
    uint64_t sdo[2] = {(uint64_t)lddo * 2, (uint64_t)N * lddo * 2};
    if (make_tmap(&tdo, dout, 2, 1, 3, ddo, sdo, box)) return -1;
    {
        const size_t smem = 2 * AB_RES + AB_STAGES * 2 * AB_CHUNK + 128 + 4 * AB_T * 4;
        VOG_CUDA(cudaFuncSetAttribute(tc_attn_bwd_sdp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int nitems = BH * p.qtiles;
        const int sms = num_sms() > 0 ? num_sms() : 148;
        tc_attn_bwd_sdp_kernel<<<nitems < sms ? nitems : sms, AB_THREADS, smem, st>>>(tq, tk, tv, tdo, p);
        if (check_launch("tc_attn_bwd_sdp")) return -1;
    }
    __nv_bfloat16* g = reinterpret_cast<__nv_bfloat16*>(dqkv);
    // dQ = dS K, dK = dS^T Q, dV = P^T dO
    if (launch_bgemm(false, dS, Npad, k, dq, sq, 0, 0, Bt, N, H, dhp, g, ldg, 0, st)) return -1;
    if (launch_bgemm(true, dS, Npad, q, dq, sq, 0, 0, Bt, N, H, dhp, g, ldg, H * dhp, st)) return -1;
    if (launch_bgemm(true, P, Npad, dout, ddo, sdo, dhp, 1, Bt, N, H, dhp, g, ldg, 2 * H * dhp, st)) return -1;
    if (bias_mode == 1) {
        const long long n = (long long)BH * N;
        attn_dbias_fold_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(drow, dcol, da, dbpe, Bt, N, H, p.nbox);
        if (check_launch("attn_dbias_fold")) return -1;
    }
    return 0;
}

}  // namespace vog
