// tcgen05 GEMM for sm_100a:  C[M,N] = epi(A[M,K] . W[N,K]^T), A and W K-major (row-major, K
// contiguous), operands bf16 (kind::f16) or tf32-rounded fp32 (kind::tf32), fp32 accumulation in
// TMEM.
//
//   persistent grid (one CTA per SM), static schedule over (128 x BN output tile, K-split) work
//   items; BN <= 256
//   warp 0      TMA producer: {128 B x 128 rows} A box + {128 B x BN rows} W box per stage,
//               128B-swizzled, ring of `stages` slots guarded by full/empty mbarriers
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (4 x K=32 B per stage),
//               tcgen05.commit releases the smem slot / publishes the accumulator
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time from one of two TMEM accumulators (so
//               the next item's MMAs overlap this item's epilogue); each warp transposes its
//               32x32 chunk through a swizzled smem patch so that global loads (residual) and
//               stores (fp32 / bf16 / tf32 copies) are 128-byte coalesced row segments
//   split-K     small-M problems (the gt5 configurations) cannot fill 148 SMs with 128-row tiles:
//               K is split across CTAs, raw fp32 partials go to a caller-provided workspace and a
//               second, elementwise kernel sums them and applies the epilogue
//
// epilogues: bias / ReLU / fp32 residual / low-precision copy / row replication, or the QKV scatter
// that writes Q,K as [Bt,H,N,dhp] bf16 and V transposed as [Bt,H,dhp,Npad] for the attention kernel.
//
// Replaces the cuBLAS calls behind nn.Linear in code/transformer_code.py:57-60,80-81,169-172,180,186
// and code/mdl_vog.py:202-207,224-230,291-314,675-677.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vog {

using namespace tc;

constexpr int GM_BM = 128;
constexpr int GM_EPI_WARPS = 8;                    // two per TMEM lane quarter: they split the column chunks
constexpr int GM_THREADS = 64 + 32 * GM_EPI_WARPS;
constexpr int GM_MAX_STAGES = 8;
constexpr int GM_A_BYTES = GM_BM * 128;
constexpr int GM_EPI_BYTES = GM_EPI_WARPS * 32 * 32 * 4;      // one 32x32 fp32 patch per epilogue warp

struct GemmParams {
    int M, N, K, BN;
    int num_k_blocks, num_m_blocks, num_n_blocks, bk_elems;
    int splits, kb_per_split, fast;
    int lq_stage;                 // bytes of the factorised-QKV language staging area (0 = read through L1)
    float* partial;               // [splits, M, N] when splits > 1
    uint32_t idesc, tmem_cols;
    int stages;
    long long* trace;             // debug: [8] clock64 stamps of CTA 0 (vog_debug_gemm_trace)
    TcEpilogue e;
};

#define GM_TRACE(i) do { if (p.trace != nullptr && blockIdx.x == 0) p.trace[i] = clock64(); } while (0)

// ---- epilogue on a coalesced float4 (4 consecutive columns of one row) ---------------------------
__device__ __forceinline__ void epi_apply_store(const TcEpilogue& e, int M, int N, int m, int n, float4 v,
                                                bool add_bias_relu)
{
    if (m >= M || n >= N) return;
    const bool full = n + 3 < N;
    float x[4] = {v.x, v.y, v.z, v.w};
    if (add_bias_relu) {
        if (e.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < N) x[j] += __ldg(e.bias + n + j);
        }
        if (e.relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.f);
        }
    }
    if (e.residual) {
        const float* rr = e.residual + (size_t)m * e.ldr + n;
        if (full && (reinterpret_cast<uintptr_t>(rr) & 15) == 0) {
            const float4 r4 = *reinterpret_cast<const float4*>(rr);
            x[0] += r4.x; x[1] += r4.y; x[2] += r4.z; x[3] += r4.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < N) x[j] += rr[j];
        }
    }
    for (int rep = 0; rep < e.rep; ++rep) {
        const size_t row = (size_t)m * e.rep + rep;
        if (e.out_f32) {
            float* dst = e.out_f32 + row * e.ldc + n;
            if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) dst[j] = x[j];
            }
        }
        if (e.out_lp) {
            if (e.lp_kind == 1) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.out_lp) + row * e.ldlp + n;
                if (full && (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                    *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) dst[j] = __float2bfloat16_rn(x[j]);
                }
            } else {
                float* dst = reinterpret_cast<float*>(e.out_lp) + row * e.ldlp + n;
                if (full && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    *reinterpret_cast<float4*>(dst) = make_float4(to_tf32(x[0]), to_tf32(x[1]), to_tf32(x[2]), to_tf32(x[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) dst[j] = to_tf32(x[j]);
                }
            }
        }
    }
}

// ---- epilogues: one per kernel instantiation so that each kernel carries only the code it runs ----
// All three read the accumulator 32 columns at a time (thread = TMEM lane = output row); the next
// chunk's tcgen05.ld is issued before the current chunk is stored so TMEM latency is hidden.
enum { EPI_FAST = 0, EPI_QKV = 1, EPI_GENERIC = 2, EPI_QKVF = 3, EPI_LIN2 = 4 };

// transpose a 32x32 fp32 chunk through the warp's swizzled smem patch: row = lane on the way in ...
__device__ __forceinline__ void patch_store(float* patch, int lane, const uint32_t (&r)[32]) {
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4)
        *reinterpret_cast<float4*>(patch + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
            make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]),
                        __uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3]));
}
// ... 4 rows x 8 float4 per instruction on the way out (128 B coalesced row segments)
__device__ __forceinline__ float4 patch_load(const float* patch, int row, int c4) {
    return *reinterpret_cast<const float4*>(patch + row * 32 + ((c4 ^ (row & 7)) << 2));
}

// EPI_FAST (host-verified: N % 32 == 0, rep == 1, every row segment 16-byte aligned): unguarded
// float4 traffic, bias / residual of the next chunk prefetched while the current one is stored - the
// first chunk even before the accumulator is ready - so their latency hides behind the MMAs
__device__ __forceinline__ void epi_fast(const GemmParams& p, const TcEpilogue& pe, float* patch, uint32_t t_acc,
                                         uint32_t tfull, uint32_t parity, int m_blk, int n_blk, int g, int half,
                                         int lane, bool trace)
{
    const int r_in = lane >> 3, c4 = lane & 7;
    const int m0 = m_blk * GM_BM + 32 * g;
    const int rows = p.M - m0;                                  // rows of this warp's slab that exist
    const size_t col = (size_t)n_blk * p.BN + c4 * 4;
    // residual row pointers of this lane's 8 rows: plain [M,N] matrix, or gathered [vis | lang] rows
    const float* rp[8];
    const bool has_res = pe.residual != nullptr || pe.res_vis != nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + i * 4 + r_in;
        rp[i] = nullptr;
        if (i * 4 + r_in < rows) {
            if (pe.residual) rp[i] = pe.residual + (size_t)m * pe.ldr + col;
            else if (pe.res_vis) {
                const int seq = pe.nsrl * pe.nppf2;
                const int bt = m / seq, rem = m - bt * seq;
                const int s_ = rem / pe.nppf2, pp = rem - s_ * pe.nppf2;
                const int cb = n_blk * p.BN;
                rp[i] = cb < pe.dv ? pe.res_vis + ((size_t)bt * pe.nppf2 + pp) * pe.ldv + cb + c4 * 4
                                   : pe.res_lang + ((size_t)(bt / pe.nfrm) * pe.nsrl + s_) * pe.ldl + (cb - pe.dv) + c4 * 4;
            }
        }
    }
    const float* biasp = pe.bias ? pe.bias + col : nullptr;
    float* o32 = pe.out_f32 ? pe.out_f32 + (size_t)(m0 + r_in) * pe.ldc + col : nullptr;
    __nv_bfloat16* obf = (pe.out_lp && pe.lp_kind == 1)
        ? reinterpret_cast<__nv_bfloat16*>(pe.out_lp) + (size_t)(m0 + r_in) * pe.ldlp + col : nullptr;
    float* otf = (pe.out_lp && pe.lp_kind != 1)
        ? reinterpret_cast<float*>(pe.out_lp) + (size_t)(m0 + r_in) * pe.ldlp + col : nullptr;
    const size_t o32_step = 4 * (size_t)pe.ldc, lp_step = 4 * (size_t)pe.ldlp;
    float4 res[8];
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) res[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prefetch = [&](int c0) {
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (rp[i]) res[i] = __ldg(reinterpret_cast<const float4*>(rp[i] + c0));
        }
        if (biasp) b4 = __ldg(reinterpret_cast<const float4*>(biasp + c0));
    };
    const int cfirst = 32 * half;                               // this warp's chunks: cfirst, cfirst + 64, ...
    if (cfirst < p.BN) prefetch(cfirst);
    mbar_wait(tfull, parity);
    tc_fence_after();
    if (trace) GM_TRACE(5);
    uint32_t r[32];
    if (cfirst < p.BN) tmem_ld32(t_acc + cfirst, r);
#pragma unroll 1
    for (int c0 = cfirst; c0 < p.BN; c0 += 64) {
        tmem_wait_ld();
        patch_store(patch, lane, r);
        __syncwarp();
        const float4 bc = b4;
        const bool more = c0 + 64 < p.BN;
        if (more && biasp) b4 = __ldg(reinterpret_cast<const float4*>(biasp + c0 + 64));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + r_in;
            float4 v = patch_load(patch, row, c4);
            const bool ok = row < rows;                        // guards single memory instructions -> predication, no branches
            v.x += bc.x; v.y += bc.y; v.z += bc.z; v.w += bc.w;
            if (pe.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            v.x += res[i].x; v.y += res[i].y; v.z += res[i].z; v.w += res[i].w;
            // this row's residual for the NEXT chunk is requested as soon as the current one is consumed
            if (has_res && more && ok) res[i] = __ldg(reinterpret_cast<const float4*>(rp[i] + c0 + 64));
            if (o32 && ok) *reinterpret_cast<float4*>(o32 + i * o32_step + c0) = v;
            if (obf && ok) *reinterpret_cast<uint2*>(obf + i * lp_step + c0) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            if (otf && ok) *reinterpret_cast<float4*>(otf + i * lp_step + c0) =
                make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        }
        if (c0 + 64 < p.BN) tmem_ld32(t_acc + c0 + 64, r);       // r is dead during the stores: fewer live registers
        __syncwarp();
    }
}

// split-K partial tile: raw fp32 accumulator rows into partial[split][M][N] (N % BN == 0, 16-byte aligned)
__device__ __forceinline__ void epi_partial(const GemmParams& p, float* out, float* patch, uint32_t t_acc,
                                            uint32_t tfull, uint32_t parity, int m_blk, int n_blk, int g, int half,
                                            int lane)
{
    const int r_in = lane >> 3, c4 = lane & 7;
    const int m0 = m_blk * GM_BM + 32 * g;
    const int rows = p.M - m0;
    float* o32 = out + (size_t)(m0 + r_in) * p.N + (size_t)n_blk * p.BN + c4 * 4;
    const size_t step = 4 * (size_t)p.N;
    const int cfirst = 32 * half;
    mbar_wait(tfull, parity);
    tc_fence_after();
    uint32_t r[32];
    if (cfirst < p.BN) tmem_ld32(t_acc + cfirst, r);
#pragma unroll 1
    for (int c0 = cfirst; c0 < p.BN; c0 += 64) {
        tmem_wait_ld();
        patch_store(patch, lane, r);
        if (c0 + 64 < p.BN) tmem_ld32(t_acc + c0 + 64, r);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = patch_load(patch, i * 4 + r_in, c4);
            if (i * 4 + r_in < rows) *reinterpret_cast<float4*>(o32 + i * step + c0) = v;
        }
        __syncwarp();
    }
}

// EPI_QKV: column block n_blk = which*H + h (BN == dhp).  Q, K and V all go through the transpose patch
// and are written as [Bt,H,N,dhp] bf16 rows (64 B row segments per lane group); the attention kernel takes
// V in this natural layout as an MN-major tcgen05 operand, so no transposed copy is ever produced
__device__ __forceinline__ void epi_qkv(const GemmParams& p, const TcEpilogue& pe, float* patch, uint32_t t_acc,
                                        uint32_t tfull, uint32_t parity, int m_blk, int n_blk, int g, int half,
                                        int lane)
{
    const int which = n_blk / pe.n_heads, h = n_blk % pe.n_heads;
    const int m0 = m_blk * GM_BM + 32 * g;
    uint32_t r[32];
    const int r_in = lane >> 3, c4 = lane & 7;
    __nv_bfloat16* base = which == 0 ? pe.q : (which == 1 ? pe.k : pe.v);
    __nv_bfloat16* dst[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + i * 4 + r_in;
        const int bt = m / pe.seq_n, ii = m % pe.seq_n;
        dst[i] = m < p.M ? base + (((size_t)bt * pe.n_heads + h) * pe.seq_n + ii) * pe.dhp + c4 * 4 : nullptr;
    }
    const int cfirst = 32 * half;
    mbar_wait(tfull, parity);
    tc_fence_after();
    if (cfirst < p.BN) tmem_ld32(t_acc + cfirst, r);
#pragma unroll 1
    for (int c0 = cfirst; c0 < p.BN; c0 += 64) {
        tmem_wait_ld();
        patch_store(patch, lane, r);
        if (c0 + 64 < p.BN) tmem_ld32(t_acc + c0 + 64, r);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = patch_load(patch, i * 4 + r_in, c4);
            const uint2 o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
            if (dst[i]) *reinterpret_cast<uint2*>(dst[i] + c0) = o;
        }
        __syncwarp();
    }
}

// EPI_QKVF: factorised QKV (TcEpilogue mode 2).  Rows are visual tokens m = bt*nppf2 + p; every row is
// written for each of the nsrl language slots with the projected language row added:
//     out(bt, s, p)[col] = acc[m, col] + lq[(bt / nfrm)*nsrl + s, col]
// The lq rows this tile needs (<= 2 queries x nsrl slots x BN columns when a query has >= 127 visual
// rows) are staged in shared memory by the four epilogue warps while the MMAs of the tile run; tiny
// configurations with more queries per tile read them through L1 instead (lqs == nullptr).
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * GM_EPI_WARPS) : "memory"); }

__device__ __forceinline__ void epi_qkvf(const GemmParams& p, const TcEpilogue& pe, float* patch, float* lqs,
                                         uint32_t t_acc, uint32_t tfull, uint32_t parity, int m_blk, int n_blk,
                                         int g, int half, int lane)
{
    const int which = n_blk / pe.n_heads, h = n_blk % pe.n_heads;
    const int m0 = m_blk * GM_BM + 32 * g;
    const int colbase = n_blk * p.BN;
    const int b0 = (m_blk * GM_BM / pe.nppf2) / pe.nfrm;          // query of the tile's first row
    if (lqs) {
        const int nq = (p.M / pe.nppf2) / pe.nfrm;
        const int bn4 = p.BN >> 2;
        const int nvec = 2 * pe.nsrl * bn4;
        epi_bar();                                                // the previous tile's readers are done
        for (int v = (half * 4 + g) * 32 + lane; v < nvec; v += 32 * GM_EPI_WARPS) {
            const int row = v / bn4, c = v - row * bn4;            // row = slot*nsrl + s
            const int b = b0 + row / pe.nsrl;
            if (b < nq)
                reinterpret_cast<float4*>(lqs)[v] = __ldg(reinterpret_cast<const float4*>(
                    pe.lq + ((size_t)b * pe.nsrl + row % pe.nsrl) * pe.ldq + colbase) + c);
        }
        epi_bar();
    }
    uint32_t r[32];
    const int r_in = lane >> 3, c4 = lane & 7;
    __nv_bfloat16* base = which == 0 ? pe.q : (which == 1 ? pe.k : pe.v);
    __nv_bfloat16* dst[8];
    const float* lqi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + i * 4 + r_in;
        const bool ok = m < p.M;
        const int bt = ok ? m / pe.nppf2 : 0, pp = ok ? m % pe.nppf2 : 0;
        const int b = bt / pe.nfrm;
        dst[i] = ok ? base + (((size_t)bt * pe.n_heads + h) * pe.seq_n + pp) * pe.dhp + c4 * 4 : nullptr;
        lqi[i] = (lqs ? lqs + (size_t)(ok ? b - b0 : 0) * pe.nsrl * p.BN
                      : pe.lq + (size_t)b * pe.nsrl * pe.ldq + colbase) + c4 * 4;
    }
    const size_t lstep = lqs ? (size_t)p.BN : (size_t)pe.ldq;
    const size_t slot_step = (size_t)pe.nppf2 * pe.dhp;
    const int cfirst = 32 * half;
    mbar_wait(tfull, parity);
    tc_fence_after();
    if (cfirst < p.BN) tmem_ld32(t_acc + cfirst, r);
#pragma unroll 1
    for (int c0 = cfirst; c0 < p.BN; c0 += 64) {
        tmem_wait_ld();
        patch_store(patch, lane, r);
        if (c0 + 64 < p.BN) tmem_ld32(t_acc + c0 + 64, r);
        __syncwarp();
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = patch_load(patch, i * 4 + r_in, c4);
#pragma unroll 1
        for (int s_ = 0; s_ < pe.nsrl; ++s_) {
            float4 l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) l[i] = *reinterpret_cast<const float4*>(lqi[i] + s_ * lstep + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint2 o = make_uint2(pack_bf16(v[i].x + l[i].x, v[i].y + l[i].y), pack_bf16(v[i].z + l[i].z, v[i].w + l[i].w));
                if (dst[i]) *reinterpret_cast<uint2*>(dst[i] + s_ * slot_step + c0) = o;
            }
        }
        __syncwarp();
    }
}

// EPI_LIN2 (TcEpilogue mode 3): the scorer tail inside the lin2[0] GEMM.  Thread = output row: it walks the
// row's BN = N accumulator columns once, h = relu(acc + bias), logit = h . w2 + b2, and writes the logit and
// the masked sigmoid score at the row's inverse-regrouped position.  Only the first warp of every TMEM lane
// quarter works (the row dot product is not split across warps); nothing of the [M, N] hidden matrix reaches HBM.
__device__ __forceinline__ void epi_lin2(const GemmParams& p, const TcEpilogue& pe, uint32_t t_acc, uint32_t tfull,
                                         uint32_t parity, int m_blk, int g, int half, int lane)
{
    mbar_wait(tfull, parity);
    tc_fence_after();
    if (half != 0) return;
    const long long m = (long long)m_blk * GM_BM + 32 * g + lane;
    float acc = 0.f;
    uint32_t r[32];
    tmem_ld32(t_acc, r);
#pragma unroll 1
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (c0 + 32 < p.BN) tmem_ld32(t_acc + c0 + 32, r);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(pe.bias + c0) + j4);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(pe.w2 + c0) + j4);
            acc = fmaf(fmaxf(v[4 * j4 + 0] + b4.x, 0.f), w4.x, acc);
            acc = fmaf(fmaxf(v[4 * j4 + 1] + b4.y, 0.f), w4.y, acc);
            acc = fmaf(fmaxf(v[4 * j4 + 2] + b4.z, 0.f), w4.z, acc);
            acc = fmaf(fmaxf(v[4 * j4 + 3] + b4.w, 0.f), w4.w, acc);
        }
    }
    if (m >= p.M) return;
    long long t = m;
    const int pp = (int)(t % pe.nppf2); t /= pe.nppf2;
    const int s_ = (int)(t % pe.nsrl); t /= pe.nsrl;
    const int f = (int)(t % pe.nfrm);
    const int b = (int)(t / pe.nfrm);
    const int P = pe.nfrm * pe.nppf2;
    const int pidx = f * pe.nppf2 + pp;                     // proposal index inside the query
    const int vid = pe.spat ? (pidx / pe.nppf) % pe.ncmp : pidx / (pe.nfrm0 * pe.nppf);
    const float logit = acc + __ldg(pe.b2);
    const size_t o = ((size_t)b * pe.nsrl + s_) * P + pidx;
    pe.logits[o] = logit;
    const float mk = (float)pe.srl_msk[(size_t)b * pe.nsrl + s_] * (float)pe.cmp_msk[(size_t)b * pe.ncmp + vid];
    pe.scores[o] = (1.f / (1.f + expf(-logit))) * mk;
}

// EPI_GENERIC: any N / alignment / row replication (guarded element-wise fallbacks inside)
__device__ __forceinline__ void epi_generic(const GemmParams& p, const TcEpilogue& pe, float* patch, uint32_t t_acc,
                                            uint32_t tfull, uint32_t parity, int m_blk, int n_blk, int g, int half,
                                            int lane)
{
    const int r_in = lane >> 3, c4 = lane & 7;
    const int m0 = m_blk * GM_BM + 32 * g;
    mbar_wait(tfull, parity);
    tc_fence_after();
    uint32_t r[32];
#pragma unroll 1
    for (int c0 = 32 * half; c0 < p.BN; c0 += 64) {
        tmem_ld32(t_acc + c0, r);
        tmem_wait_ld();
        patch_store(patch, lane, r);
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            const int row = i * 4 + r_in;
            const float4 v = patch_load(patch, row, c4);
            epi_apply_store(pe, p.M, p.N, m0 + row, n_blk * p.BN + c0 + c4 * 4, v, true);
        }
        __syncwarp();
    }
}

// ---- kernel -----------------------------------------------------------------------------------
template <bool kTF32, int kEpi>
__global__ void __launch_bounds__(GM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const GemmParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);
    const uint32_t stage_bytes = GM_A_BYTES + p.BN * 128;
    const uint32_t epi_off = p.stages * stage_bytes;
    const uint32_t bar_off = epi_off + GM_EPI_BYTES + p.lq_stage;
    const uint32_t bar_base = smem_base + bar_off;
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(
        smem_gen + bar_off + 8 * (2 * GM_MAX_STAGES + 4));
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (GM_MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * GM_MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * GM_MAX_STAGES + 2 + a); };
    auto a_smem = [&](int s) { return smem_base + s * stage_bytes; };
    auto b_smem = [&](int s) { return smem_base + s * stage_bytes + GM_A_BYTES; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) GM_TRACE(0);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), GM_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr)), p.tmem_cols);
    pdl_trigger();                                   // the next kernel of the stream may set itself up behind this one
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                                      // everything above overlapped the previous kernel's tail
    if (threadIdx.x == 0) GM_TRACE(1);
    const int nitems = p.num_m_blocks * p.num_n_blocks * p.splits;
    // item -> (m_blk, n_blk, split): splits innermost so that the CTAs of one wave share A/W tiles in L2
    auto decode = [&](int item, int& m_blk, int& n_blk, int& kb0, int& kb1) {
        const int split = item % p.splits;
        const int tile = item / p.splits;
        m_blk = tile / p.num_n_blocks;
        n_blk = tile % p.num_n_blocks;
        kb0 = split * p.kb_per_split;
        kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
        return split;
    };

    // Producer and MMA issuer run their loops with the WHOLE warp (warp-uniform control flow) and elect one
    // lane per TMA / tcgen05 instruction: under a divergent `if (lane == 0)` every uniform-datapath
    // instruction is wrapped in an ELECT / BRA.U.ANY loop (~11 SASS instructions per tcgen05.mma).
    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            decode(item, m_blk, n_blk, kb0, kb1);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(s), stage_bytes);
                    tma_load_2d(a_smem(s), &tma_a, full_bar(s), kb * p.bk_elems, m_blk * GM_BM);
                    tma_load_2d(b_smem(s), &tma_b, full_bar(s), kb * p.bk_elems, n_blk * p.BN);
                    if (item == (int)blockIdx.x && kb == kb0) GM_TRACE(2);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0; int acc = 0; uint32_t acc_ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            decode(item, m_blk, n_blk, kb0, kb1);
            mbar_wait(tempty_bar(acc), acc_ph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * p.BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                if (elect_one()) {
                    if (item == (int)blockIdx.x && kb == kb0) GM_TRACE(3);
                    const uint64_t ad = umma_desc_sw128(a_smem(s));
                    const uint64_t bd = umma_desc_sw128(b_smem(s));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma<kTF32>(d_tmem, ad + 2 * k, bd + 2 * k, p.idesc, (kb > kb0) || (k != 0));
                    umma_commit(empty_bar(s));
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            if (elect_one()) {
                umma_commit(tfull_bar(acc));
                if (item == (int)blockIdx.x) GM_TRACE(4);
            }
            __syncwarp();
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1;
        }
    } else {
        const int g = warp & 3;                    // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;          // which of the quarter's two warps: even / odd 32-column chunks
        float* patch = reinterpret_cast<float*>(smem_gen + epi_off) + (warp - 2) * 1024;     // [32 rows][8 float4]
        float* lqs = p.lq_stage ? reinterpret_cast<float*>(smem_gen + epi_off + GM_EPI_BYTES) : nullptr;
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int m_blk, n_blk, kb0, kb1;
            const int split = decode(item, m_blk, n_blk, kb0, kb1);
            const uint32_t t_acc = tmem_base + ((uint32_t)(32 * g) << 16) + acc * p.BN;
            const bool trace = item == (int)blockIdx.x && threadIdx.x == 64;
            const uint32_t tf = tfull_bar(acc);
            if (p.splits > 1) {                    // split-K: raw partials, no epilogue math
                float* part = p.partial + (size_t)split * p.M * p.N;
                if constexpr (kEpi == EPI_FAST) {
                    if (trace) GM_TRACE(5);
                    epi_partial(p, part, patch, t_acc, tf, acc_ph, m_blk, n_blk, g, half, lane);
                } else {
                    TcEpilogue pe;
                    pe.out_f32 = part;
                    pe.ldc = p.N;
                    epi_generic(p, pe, patch, t_acc, tf, acc_ph, m_blk, n_blk, g, half, lane);
                }
            } else if constexpr (kEpi == EPI_FAST) {
                epi_fast(p, p.e, patch, t_acc, tf, acc_ph, m_blk, n_blk, g, half, lane, trace);
            } else if constexpr (kEpi == EPI_QKV) {
                epi_qkv(p, p.e, patch, t_acc, tf, acc_ph, m_blk, n_blk, g, half, lane);
            } else if constexpr (kEpi == EPI_QKVF) {
                epi_qkvf(p, p.e, patch, lqs, t_acc, tf, acc_ph, m_blk, n_blk, g, half, lane);
            } else if constexpr (kEpi == EPI_LIN2) {
                epi_lin2(p, p.e, t_acc, tf, acc_ph, m_blk, g, half, lane);
            } else {
                epi_generic(p, p.e, patch, t_acc, tf, acc_ph, m_blk, n_blk, g, half, lane);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (trace) GM_TRACE(6);
            acc ^= 1;
            if (acc == 0) acc_ph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) GM_TRACE(7);
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

static thread_local long long* g_gemm_trace = nullptr;
static thread_local int g_reserved_sms = 0;
void tc_gemm_set_reserved_sms(int n) { g_reserved_sms = n > 0 ? n : 0; }
void tc_gemm_set_trace(long long* buf) { g_gemm_trace = buf; }

// sum of split-K partials + epilogue, one float4 per thread
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, TcEpilogue e)
{
    const int n4 = (N + 3) / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= (long long)M * n4) return;
    const int m = (int)(idx / n4), n = (int)(idx % n4) * 4;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = (N % 4) == 0;
    for (int s = 0; s < splits; ++s) {
        const float* src = partial + ((size_t)s * M + m) * N + n;
        if (vec) {
            const float4 v = *reinterpret_cast<const float4*>(src);
            x[0] += v.x; x[1] += v.y; x[2] += v.z; x[3] += v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < N) x[j] += src[j];
        }
    }
    epi_apply_store(e, M, N, m, n, make_float4(x[0], x[1], x[2], x[3]), true);
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

// K-split factor: only when the tile grid leaves most SMs idle and K is deep enough to be worth it
int tc_gemm_splits(int M, int N, int K, int tf32, int BN, int qkv_mode)
{
    if (qkv_mode) return 1;                  // also: gathered residual (no split-K epilogue for it)
    const int bk = tf32 ? 32 : 64;
    const int nkb = cdiv(K, bk);
    const int tiles = cdiv(M, GM_BM) * cdiv(N, BN);
    const int sms = num_sms() > 0 ? num_sms() : 148;
    if (tiles * 2 > sms || nkb < 16) return 1;          // shallow K: a second (reduction) launch costs more than it saves
    int s = sms / tiles;
    if (s > nkb / 4) s = nkb / 4;            // at least 4 k-blocks per split
    if (s > 32) s = 32;
    return s < 2 ? 1 : s;
}

long long tc_gemm_workspace_bytes(int M, int N, int K, int tf32, int BN)
{
    const int s = tc_gemm_splits(M, N, K, tf32, BN, 0);
    return s > 1 ? (long long)s * M * N * 4 : 0;
}

int tc_gemm(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, int tf32,
            int BN, const TcEpilogue& epi, void* workspace, long long workspace_bytes, cudaStream_t st)
{
    if (M == 0 || N == 0) return 0;
    const int eb = tf32 ? 4 : 2;
    const int bk = 128 / eb;
    VOG_REQUIRE(BN >= 32 && BN <= 256 && BN % 32 == 0, "tc_gemm: BN=%d must be a multiple of 32 in [32,256]", BN);
    VOG_REQUIRE(K > 0 && (K * eb) % 16 == 0, "tc_gemm: K=%d rows must be 16-byte multiples", K);
    VOG_REQUIRE(lda >= K && ldw >= K, "tc_gemm: bad leading dimension");
    VOG_REQUIRE(epi.rep >= 1, "tc_gemm: rep must be >= 1");
    if (epi.mode == 1 || epi.mode == 2) {
        VOG_REQUIRE(BN == epi.dhp && N == 3 * epi.n_heads * epi.dhp, "tc_gemm: qkv epilogue needs BN == dhp, N == 3*H*dhp");
        VOG_REQUIRE(epi.q && epi.k && epi.v && epi.seq_n > 0, "tc_gemm: bad qkv epilogue");
        if (epi.mode == 2) {
            VOG_REQUIRE(epi.lq && epi.nsrl > 0 && epi.nppf2 > 0 && epi.nfrm > 0 && epi.seq_n == epi.nsrl * epi.nppf2 &&
                        M % epi.nppf2 == 0 && (M / epi.nppf2) % epi.nfrm == 0,
                        "tc_gemm: bad factorised-qkv geometry");
            VOG_REQUIRE(epi.ldq >= N && epi.ldq % 4 == 0 && (reinterpret_cast<uintptr_t>(epi.lq) & 15) == 0,
                        "tc_gemm: language projection must be 16-byte aligned rows");
        }
    } else if (epi.mode == 3) {
        VOG_REQUIRE(BN == N && N <= 256 && N % 32 == 0, "tc_gemm: fused scorer tail needs one column tile (N=%d, BN=%d)", N, BN);
        VOG_REQUIRE(epi.bias && epi.w2 && epi.b2 && epi.srl_msk && epi.cmp_msk && epi.logits && epi.scores,
                    "tc_gemm: fused scorer tail: null operand");
        VOG_REQUIRE((reinterpret_cast<uintptr_t>(epi.bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(epi.w2) & 15) == 0,
                    "tc_gemm: fused scorer tail: bias / w2 must be 16-byte aligned");
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
    p.splits = tc_gemm_splits(M, N, K, tf32, BN, epi.mode != 0 || epi.res_vis != nullptr);
    if (p.splits > 1 && (workspace == nullptr || workspace_bytes < (long long)p.splits * M * N * 4))
        p.splits = 1;                          // caller gave no workspace: plain schedule
    p.kb_per_split = cdiv(p.num_k_blocks, p.splits);
    p.splits = cdiv(p.num_k_blocks, p.kb_per_split);
    p.partial = reinterpret_cast<float*>(workspace);
    p.idesc = umma_idesc(tf32 ? FMT_TF32 : FMT_BF16, GM_BM, BN);
    p.tmem_cols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    const int stage_bytes = GM_A_BYTES + BN * 128;
    p.lq_stage = 0;
    if (epi.mode == 2 && epi.nfrm * epi.nppf2 >= GM_BM - 1 && epi.nsrl <= 16)
        p.lq_stage = 2 * epi.nsrl * BN * 4;       // a 128-row tile then spans at most two queries
    const int budget = 227 * 1024 - 1024 /*align*/ - 256 /*barriers*/ - GM_EPI_BYTES - p.lq_stage;
    int stages = budget / stage_bytes;
    if (stages > GM_MAX_STAGES) stages = GM_MAX_STAGES;
    VOG_REQUIRE(stages >= 2, "tc_gemm: tile does not fit shared memory");
    p.stages = stages;
    p.trace = g_gemm_trace;
    p.e = epi;
    {   // unguarded float4 epilogue when every row segment the kernel touches is 16-byte aligned
        auto al = [](const void* q, int a) { return (reinterpret_cast<uintptr_t>(q) % a) == 0; };
        bool ok = (N % BN == 0) && epi.rep == 1 && epi.mode == 0;     // modes 1-3 have their own epilogues
        if (p.splits > 1) ok = (N % BN == 0) && epi.mode == 0 && al(workspace, 16);
        else {
            ok = ok && (!epi.bias || al(epi.bias, 16));
            ok = ok && (!epi.residual || (al(epi.residual, 16) && epi.ldr % 4 == 0));
            if (epi.res_vis)
                ok = ok && al(epi.res_vis, 16) && al(epi.res_lang, 16) && epi.ldv % 4 == 0 && epi.ldl % 4 == 0 &&
                     epi.dv % BN == 0;
            ok = ok && (!epi.out_f32 || (al(epi.out_f32, 16) && epi.ldc % 4 == 0));
            if (epi.out_lp) ok = ok && (epi.lp_kind == 1 ? (al(epi.out_lp, 8) && epi.ldlp % 4 == 0)
                                                         : (al(epi.out_lp, 16) && epi.ldlp % 4 == 0));
        }
        p.fast = ok ? 1 : 0;
    }
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + GM_EPI_BYTES + p.lq_stage;
    const int nitems = p.num_m_blocks * p.num_n_blocks * p.splits;
    // vog_set_reserved_sms: SMs left to a concurrently running kernel of another branch (the persistent tile
    // schedule is static, so CTAs that would have to wait for those SMs double the kernel's time)
    int avail = num_sms() - g_reserved_sms;
    if (avail < 1) avail = 1;
    const int grid = nitems < avail ? nitems : avail;
    const int epi_kind = epi.mode == 1 ? EPI_QKV : epi.mode == 2 ? EPI_QKVF : epi.mode == 3 ? EPI_LIN2
                         : (p.fast ? EPI_FAST : EPI_GENERIC);
    VOG_REQUIRE(!epi.res_vis || epi_kind == EPI_FAST, "tc_gemm: gathered residual needs the aligned fast epilogue "
                "(N %% BN == 0, dv %% BN == 0, 16-byte aligned rows)");
#define VOG_GEMM_LAUNCH(TF, EP)                                                                                   \
    do {                                                                                                          \
        VOG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<TF, EP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        VOG_CUDA(launch_pdl(tc_gemm_kernel<TF, EP>, dim3(grid), dim3(GM_THREADS), smem, st, ta, tb, p));                  \
    } while (0)
    if (tf32) {
        if (epi_kind == EPI_FAST) VOG_GEMM_LAUNCH(true, EPI_FAST);
        else if (epi_kind == EPI_QKV) VOG_GEMM_LAUNCH(true, EPI_QKV);
        else if (epi_kind == EPI_QKVF) VOG_GEMM_LAUNCH(true, EPI_QKVF);
        else if (epi_kind == EPI_LIN2) VOG_GEMM_LAUNCH(true, EPI_LIN2);
        else VOG_GEMM_LAUNCH(true, EPI_GENERIC);
    } else {
        if (epi_kind == EPI_FAST) VOG_GEMM_LAUNCH(false, EPI_FAST);
        else if (epi_kind == EPI_QKV) VOG_GEMM_LAUNCH(false, EPI_QKV);
        else if (epi_kind == EPI_QKVF) VOG_GEMM_LAUNCH(false, EPI_QKVF);
        else if (epi_kind == EPI_LIN2) VOG_GEMM_LAUNCH(false, EPI_LIN2);
        else VOG_GEMM_LAUNCH(false, EPI_GENERIC);
    }
#undef VOG_GEMM_LAUNCH
    if (check_launch("tc_gemm")) return -1;
    if (p.splits > 1) {
        const long long n = (long long)M * ((N + 3) / 4);
        VOG_CUDA(launch_pdl(splitk_reduce_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const float*)p.partial, p.splits, M, N, epi));
        return check_launch("splitk_reduce");
    }
    return 0;
}

}  // namespace vog
