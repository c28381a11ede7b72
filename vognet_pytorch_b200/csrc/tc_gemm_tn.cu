// tcgen05 "TN" GEMM for sm_100a - the weight-gradient contraction of the training step:
//
//     C[N1, N2] += sum_m A[m, N1] * B[m, N2]          (dW = dY^T . X over all rows m of the activation matrices)
//
// A = dY [Mrows, N1] and B = X [Mrows, N2] are both stored row-major (the way the forward / backward kernels
// leave them), i.e. the contraction index m is the SLOW index of both operands.  Neither is transposed in HBM:
// TMA loads {64 columns x 64 rows} boxes (128 B x 64, 128B swizzle) straight from the natural layout and the MMA
// consumes them as MN-MAJOR shared-memory operands (instruction-descriptor bits 15 / 16, the same form the
// attention kernel uses for V), so every activation byte is read exactly once per output tile column / row.
//
//   persistent grid, work items = (128 x BN output tile, K-split); each split reduces a slab of rows m and adds
//   its fp32 tile into C with vector red.global.add (C holds the running gradient: zero-initialised by the
//   caller, or an accumulation over micro-batches)
//   warp 0 TMA producer, warp 1 MMA issuer (4 x K=16 per 64-row stage), warps 2-9 epilogue (two TMEM
//   accumulators: the next item's MMAs overlap this item's reduction)
//
// Replaces what torch autograd does for every nn.Linear weight of the path (mm of the transposed output
// gradient with the saved input): code/transformer_code.py:57-60,80-81,169-172,180,186; code/mdl_vog.py:202-207,
// 224-230 - as called from utils/trn_utils.py:504 (loss.backward()).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vog {

using namespace tc;

constexpr int TN_BM = 128;
constexpr int TN_BK = 64;                          // rows m per pipeline stage
constexpr int TN_EPI_WARPS = 8;
constexpr int TN_THREADS = 64 + 32 * TN_EPI_WARPS;
constexpr int TN_MAX_STAGES = 8;
constexpr int TN_SUB = TN_BK * 128;                // one [64 rows x 128 B] sub-tile

struct GemmTnParams {
    int N1, N2, K, BN;
    int num_k_blocks, num_m_blocks, num_n_blocks;
    int splits, kb_per_split, stages;
    uint32_t idesc, tmem_cols;
    float* C; long long ldc;
};

__global__ void __launch_bounds__(TN_THREADS, 1)
tc_gemm_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const GemmTnParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const int nsub_b = p.BN / 64;
    const uint32_t a_bytes = 2 * TN_SUB, b_bytes = nsub_b * TN_SUB;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const uint32_t bar_off = p.stages * stage_bytes;
    const uint32_t bar_base = smem_base + bar_off;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * TN_MAX_STAGES + 4));
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TN_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TN_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TN_MAX_STAGES + 2 + a); };
    auto a_smem = [&](int s) { return smem_base + s * stage_bytes; };
    auto b_smem = [&](int s) { return smem_base + s * stage_bytes + a_bytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TN_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int nitems = p.num_m_blocks * p.num_n_blocks * p.splits;
    auto decode = [&](int item, int& m_blk, int& n_blk, int& kb0, int& kb1) {
        const int split = item % p.splits;
        const int tile = item / p.splits;
        m_blk = tile / p.num_n_blocks;
        n_blk = tile % p.num_n_blocks;
        kb0 = split * p.kb_per_split;
        kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
    };

    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            decode(item, m_blk, n_blk, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                    // rows beyond K and columns beyond N1 / N2 are zero-filled by TMA: they add nothing
                    for (int i = 0; i < 2; ++i)
                        tma_load_2d(a_smem(s) + i * TN_SUB, &tma_a, full_bar(s), m_blk * TN_BM + i * 64, kb * TN_BK);
                    for (int i = 0; i < nsub_b; ++i)
                        tma_load_2d(b_smem(s) + i * TN_SUB, &tma_b, full_bar(s), n_blk * p.BN + i * 64, kb * TN_BK);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0;
        // MN-major 128B-swizzled operands: 8-row (K) groups 1024 B apart (SBO), 64-element (MN) groups one sub-tile apart
        // (LBO); a K = 16 step advances the start address by 16 rows x 128 B
        const uint32_t lbo = ((uint32_t)TN_SUB >> 4) << 16;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            decode(item, m_blk, n_blk, kb0, kb1);
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * p.BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t a_lo = ((a_smem(s) >> 4) & 0x3FFF) | lbo;
                const uint32_t b_lo = ((b_smem(s) >> 4) & 0x3FFF) | lbo;
                if (elect_one()) {
                    if (kb == kb0) umma_bf16_lo<false>(d_tmem, a_lo, b_lo, p.idesc);
                    else umma_bf16_lo<true>(d_tmem, a_lo, b_lo, p.idesc);
#pragma unroll
                    for (int k = 1; k < TN_BK / 16; ++k)
                        umma_bf16_lo<true>(d_tmem, a_lo + k * (2048 >> 4), b_lo + k * (2048 >> 4), p.idesc);
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
        const int g = warp & 3;
        const int half = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            decode(item, m_blk, n_blk, kb0, kb1);
            const uint32_t t_acc = tmem_base + ((uint32_t)(32 * g) << 16) + acc * p.BN;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            const int m = m_blk * TN_BM + 32 * g + lane;               // output row = TMEM lane
            float* crow = p.C + (long long)m * p.ldc + (long long)n_blk * p.BN;
            uint32_t r[32];
#pragma unroll 1
            for (int c0 = 32 * half; c0 < p.BN; c0 += 64) {
                tmem_ld32(t_acc + c0, r);
                tmem_wait_ld();
                if (m < p.N1) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int n = n_blk * p.BN + c0 + j;
                        if (n + 3 < p.N2) {
                            atomicAdd(reinterpret_cast<float4*>(crow + c0 + j),
                                      make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                  __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                        } else {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (n + q < p.N2) atomicAdd(crow + c0 + j + q, __uint_as_float(r[j + q]));
                        }
                    }
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

// C[N1,N2] (fp32, ldc) += A[K,N1]^T . B[K,N2], A / B bf16 row-major with leading dimensions lda / ldb (elements)
int tc_gemm_tn(const void* A, long long lda, const void* B, long long ldb, int K, int N1, int N2, float* C,
               long long ldc, cudaStream_t st)
{
    if (N1 == 0 || N2 == 0 || K == 0) return 0;
    VOG_REQUIRE(lda >= N1 && ldb >= N2 && ldc >= N2, "tc_gemm_tn: bad leading dimension");
    VOG_REQUIRE((lda * 2) % 16 == 0 && (ldb * 2) % 16 == 0, "tc_gemm_tn: operand rows must be 16-byte multiples");
    VOG_REQUIRE(ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "tc_gemm_tn: C must be 16-byte aligned rows");
    int BN = 256;
    if (N2 <= 64) BN = 64; else if (N2 <= 128) BN = 128; else if (N2 <= 192) BN = 192;
    else if (N2 % 256 != 0 && N2 % 192 == 0) BN = 192;
    CUtensorMap ta, tb;
    uint64_t da[2] = {(uint64_t)N1, (uint64_t)K}, sa[1] = {(uint64_t)lda * 2};
    uint64_t db[2] = {(uint64_t)N2, (uint64_t)K}, sb[1] = {(uint64_t)ldb * 2};
    uint32_t box[2] = {64, (uint32_t)TN_BK};
    if (make_tmap(&ta, A, 2, 1, 2, da, sa, box)) return -1;
    if (make_tmap(&tb, B, 2, 1, 2, db, sb, box)) return -1;
    GemmTnParams p;
    p.N1 = N1; p.N2 = N2; p.K = K; p.BN = BN; p.C = C; p.ldc = ldc;
    p.num_k_blocks = cdiv(K, TN_BK);
    p.num_m_blocks = cdiv(N1, TN_BM);
    p.num_n_blocks = cdiv(N2, BN);
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    const int sms = num_sms() > 0 ? num_sms() : 148;
    int splits = (2 * sms + tiles - 1) / tiles;             // about two work items per SM
    if (splits > p.num_k_blocks / 4) splits = p.num_k_blocks / 4;
    if (splits < 1) splits = 1;
    p.kb_per_split = cdiv(p.num_k_blocks, splits);
    p.splits = cdiv(p.num_k_blocks, p.kb_per_split);
    // A and B both MN-major (transpose bits 15 and 16)
    p.idesc = umma_idesc(FMT_BF16, TN_BM, BN) | (1u << 15) | (1u << 16);
    p.tmem_cols = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    const int stage_bytes = (2 + BN / 64) * TN_SUB;
    int stages = (227 * 1024 - 1024 - 256) / stage_bytes;
    if (stages > TN_MAX_STAGES) stages = TN_MAX_STAGES;
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    const int nitems = tiles * p.splits;
    const int grid = nitems < sms ? nitems : sms;
    VOG_CUDA(cudaFuncSetAttribute(tc_gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_gemm_tn_kernel<<<grid, TN_THREADS, smem, st>>>(ta, tb, p);
    return check_launch("tc_gemm_tn");
}

}  // namespace vog
