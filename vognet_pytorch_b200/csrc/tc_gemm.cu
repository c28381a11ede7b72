// tcgen05 GEMM for sm_100a:  C[M,N] = epi(A[M,K] . W[N,K]^T), A and W K-major (row-major, K
// contiguous), operands bf16 (kind::f16) or tf32-rounded fp32 (kind::tf32), fp32 accumulation in
// TMEM.
//
//   persistent grid (one CTA per SM), static tile schedule, 128 x BN output tiles (BN <= 256)
//   warp 0      TMA producer: {128 B x 128 rows} A box + {128 B x BN rows} W box per stage,
//               128B-swizzled, ring of `stages` slots guarded by full/empty mbarriers
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=32 B per stage),
//               tcgen05.commit releases the smem slot / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time from one of two TMEM accumulators
//               (so the next tile's MMAs overlap this tile's epilogue), fused bias / ReLU /
//               residual / low-precision copy / row replication, or the QKV scatter that writes
//               Q,K as [Bt,H,N,dhp] bf16 and V transposed as [Bt,H,dhp,Npad] for the attention kernel
//
// Replaces the cuBLAS calls behind nn.Linear in code/transformer_code.py:57-60,80-81,169-172,180,186
// and code/mdl_vog.py:202-207,224-230,291-314,675-677.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vog {

using namespace tc;

constexpr int GM_BM = 128;
constexpr int GM_THREADS = 192;
constexpr int GM_MAX_STAGES = 8;
constexpr int GM_A_BYTES = GM_BM * 128;

struct GemmParams {
    int M, N, K, BN;
    int num_k_blocks, num_m_blocks, num_n_blocks, bk_elems;
    uint32_t idesc, tmem_cols;
    int stages;
    TcEpilogue e;
};

// ---- epilogues --------------------------------------------------------------------------------
__device__ __forceinline__ void store_row_f32(float* dst, const float (&v)[32], int nvalid) {
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) dst[j] = v[j];
    }
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&v)[32], int nvalid) {
    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(dst + j) =
                make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                           pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) dst[j] = __float2bfloat16_rn(v[j]);
    }
}

__device__ __forceinline__ void epilogue_std(const GemmParams& p, int m, int n0, const uint32_t (&r)[32]) {
    const TcEpilogue& e = p.e;
    const int nvalid = min(32, p.N - n0);
    if (m >= p.M || nvalid <= 0) return;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (e.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += __ldg(e.bias + n0 + j);
    }
    if (e.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (e.residual) {
        const float* rr = e.residual + (size_t)m * e.ldr + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += rr[j];
    }
    for (int rep = 0; rep < e.rep; ++rep) {
        const size_t row = (size_t)m * e.rep + rep;
        if (e.out_f32) store_row_f32(e.out_f32 + row * e.ldc + n0, v, nvalid);
        if (e.out_lp) {
            if (e.lp_kind == 1) {
                store_row_bf16(reinterpret_cast<__nv_bfloat16*>(e.out_lp) + row * e.ldlp + n0, v, nvalid);
            } else {
                float t[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) t[j] = to_tf32(v[j]);
                store_row_f32(reinterpret_cast<float*>(e.out_lp) + row * e.ldlp + n0, t, nvalid);
            }
        }
    }
}

// QKV scatter: N = 3*H*dhp, BN == dhp, tile column block n_blk = which*H + h
__device__ __forceinline__ void epilogue_qkv(const GemmParams& p, int m, int n_blk, int c0,
                                             const uint32_t (&r)[32]) {
    const TcEpilogue& e = p.e;
    if (m >= p.M) return;
    const int which = n_blk / e.n_heads, h = n_blk % e.n_heads;
    const int bt = m / e.seq_n, i = m % e.seq_n;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (which < 2) {
        __nv_bfloat16* dst = (which == 0 ? e.q : e.k) +
                             (((size_t)bt * e.n_heads + h) * e.seq_n + i) * e.dhp + c0;
        store_row_bf16(dst, v, 32);
    } else {
        __nv_bfloat16* dst = e.vt + (((size_t)bt * e.n_heads + h) * e.dhp + c0) * e.npad + i;
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[(size_t)j * e.npad] = __float2bfloat16_rn(v[j]);
    }
}

// ---- kernel -----------------------------------------------------------------------------------
template <bool kTF32>
__global__ void __launch_bounds__(GM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t stage_bytes = GM_A_BYTES + p.BN * 128;
    const uint32_t bar_base = smem_base + p.stages * stage_bytes;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_gen + p.stages * stage_bytes + 8 * (2 * GM_MAX_STAGES + 4));
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (GM_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * GM_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * GM_MAX_STAGES + 2 + a); };
    auto a_smem = [&](int s) { return smem_base + s * stage_bytes; };
    auto b_smem = [&](int s) { return smem_base + s * stage_bytes + GM_A_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int ntiles = p.num_m_blocks * p.num_n_blocks;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1);
                    mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                    tma_load_2d(a_smem(s), &tma_a, full_bar(s), kb * p.bk_elems, m_blk * GM_BM);
                    tma_load_2d(b_smem(s), &tma_b, full_bar(s), kb * p.bk_elems, n_blk * p.BN);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * p.BN;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(a_smem(s));
                    const uint64_t bd = umma_desc_sw128(b_smem(s));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma<kTF32>(d_tmem, ad + 2 * k, bd + 2 * k, p.idesc, (kb | k) != 0);
                    umma_commit(empty_bar(s));
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar(acc));
                acc ^= 1;
                if (acc == 0) acc_ph ^= 1;
            }
        }
        __syncwarp();
    } else {
        const int g = warp & 3;                    // TMEM lane quarter this warp may access
        int acc = 0; uint32_t acc_ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int m_blk = tile / p.num_n_blocks, n_blk = tile % p.num_n_blocks;
            mbar_wait(tfull_bar(acc), acc_ph);
            tc_fence_after();
            const int m = m_blk * GM_BM + 32 * g + lane;
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(32 * g) << 16) + acc * p.BN + c0, r);
                tmem_wait_ld();
                if (p.e.mode == 1) epilogue_qkv(p, m, n_blk, c0, r);
                else epilogue_std(p, m, n_blk * p.BN + c0, r);
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

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int is_bf16, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box)
{
    EncodeTiledFn enc = get_encoder();
    VOG_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    VOG_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    cuuint64_t gd[5]; cuuint64_t gs[5]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) {
        VOG_REQUIRE(strides_bytes[i] % 16 == 0, "TMA global stride %llu not a multiple of 16 bytes",
                    (unsigned long long)strides_bytes[i]);
        gs[i] = strides_bytes[i];
    }
    VOG_REQUIRE((int)box[0] * elem_bytes == 128, "TMA box inner extent must be 128 bytes");
    CUresult r = enc(out, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VOG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

int num_sms()
{
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

int tc_gemm(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int tf32,
            int BN, const TcEpilogue& epi, cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    const int eb = tf32 ? 4 : 2;
    const int bk = 128 / eb;
    VOG_REQUIRE(BN >= 32 && BN <= 256 && BN % 32 == 0, "tc_gemm: BN=%d must be a multiple of 32 in [32,256]", BN);
    VOG_REQUIRE(K > 0 && (K * eb) % 16 == 0, "tc_gemm: K=%d rows must be 16-byte multiples", K);
    VOG_REQUIRE(lda >= K && ldw >= K, "tc_gemm: bad leading dimension");
    VOG_REQUIRE(epi.rep >= 1, "tc_gemm: rep must be >= 1");
    if (epi.mode == 1) {
        VOG_REQUIRE(BN == epi.dhp && N == 3 * epi.n_heads * epi.dhp, "tc_gemm: qkv epilogue needs BN == dhp, N == 3*H*dhp");
        VOG_REQUIRE(epi.q && epi.k && epi.vt && epi.seq_n > 0 && epi.npad >= epi.seq_n, "tc_gemm: bad qkv epilogue");
    } else {
        VOG_REQUIRE(epi.out_f32 || epi.out_lp, "tc_gemm: no output");
    }
    CUtensorMap ta, tb;
    uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, sa[1] = {(uint64_t)lda * eb};
    uint32_t ba[2] = {(uint32_t)bk, (uint32_t)GM_BM};
    if (make_tmap(&ta, A, eb, !tf32, 2, da, sa, ba)) return -1;
    uint64_t db[2] = {(uint64_t)K, (uint64_t)N}, sb[1] = {(uint64_t)ldw * eb};
    uint32_t bb[2] = {(uint32_t)bk, (uint32_t)BN};
    if (make_tmap(&tb, W, eb, !tf32, 2, db, sb, bb)) return -1;

    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.BN = BN; p.bk_elems = bk;
    p.num_k_blocks = cdiv(K, bk);
    p.num_m_blocks = cdiv(M, GM_BM);
    p.num_n_blocks = cdiv(N, BN);
    p.idesc = umma_idesc(tf32 ? FMT_TF32 : FMT_BF16, GM_BM, BN);
    p.tmem_cols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    const int stage_bytes = GM_A_BYTES + BN * 128;
    const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/;
    int stages = budget / stage_bytes;
    if (stages > GM_MAX_STAGES) stages = GM_MAX_STAGES;
    VOG_REQUIRE(stages >= 2, "tc_gemm: tile does not fit shared memory");
    p.stages = stages;
    p.e = epi;
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    const int ntiles = p.num_m_blocks * p.num_n_blocks;
    const int grid = ntiles < num_sms() ? ntiles : num_sms();
    if (tf32) {
        VOG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_gemm_kernel<true><<<grid, GM_THREADS, smem, st>>>(ta, tb, p);
    } else {
        VOG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_gemm_kernel<false><<<grid, GM_THREADS, smem, st>>>(ta, tb, p);
    }
    return check_launch("tc_gemm");
}

}  // namespace vog
