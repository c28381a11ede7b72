// Glue of the fusion path as two bandwidth-bound kernels instead of ~12 library launches:
//
//  build_xmul   the multimodal transformer's input.  Token (b, f, s, p') of per-frame sequence
//               (b,f) is [vis[b, f*nppf'+p'] | lang[b,s]] (code/mdl_vog.py:316-344 concat,
//               :693-699 per-frame regroup): written once as fp32 (residual stream) + low-precision
//               copy (GEMM operand), replacing expand/cat/transpose/contiguous/cast.
//  lin2_tail    the scorer's last layer and everything after it: logit = h . w2 + b2
//               (code/mdl_vog.py:675-677), inverse regroup to [B,1,nsrl,P] (:724-737),
//               mdl_outs_eval = sigmoid(logit) * srl_arg_inds_msk * num_cmp_msk
//               (code/mdl_conc_single.py:39-48,118-122,144-154).
#include "common.cuh"
#include "kernels.h"

namespace vog {

__global__ void __launch_bounds__(256)
build_xmul_kernel(const float* __restrict__ vis, const float* __restrict__ lang, float* __restrict__ out,
                  void* __restrict__ out_lp, int lp_kind, int B, int nfrm, int nsrl, int nppf2, int dv, int dl)
{
    const int d4 = (dv + dl) / 4;
    const long long total = (long long)B * nfrm * nsrl * nppf2 * d4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % d4) * 4;
    long long m = idx / d4;
    const int pp = (int)(m % nppf2); m /= nppf2;
    const int s = (int)(m % nsrl); m /= nsrl;
    const int f = (int)(m % nfrm);
    const int b = (int)(m / nfrm);
    float4 v;
    if (c < dv) v = *reinterpret_cast<const float4*>(vis + ((size_t)b * nfrm * nppf2 + (size_t)f * nppf2 + pp) * dv + c);
    else v = *reinterpret_cast<const float4*>(lang + ((size_t)b * nsrl + s) * dl + (c - dv));
    const size_t o = (size_t)(idx / d4) * (dv + dl) + c;
    if (out) *reinterpret_cast<float4*>(out + o) = v;
    if (out_lp) {
        if (lp_kind == 1) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out_lp) + o) =
                make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(out_lp) + o) =
                make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        }
    }
}

int build_xmul(const float* vis, const float* lang, float* out, void* out_lp, int lp_kind, int B, int nfrm,
               int nsrl, int nppf2, int dv, int dl, cudaStream_t st)
{
    VOG_REQUIRE(dv % 4 == 0 && dl % 4 == 0, "build_xmul: feature dims must be multiples of 4");
    const long long total = (long long)B * nfrm * nsrl * nppf2 * ((dv + dl) / 4);
    if (total == 0) return 0;
    build_xmul_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(vis, lang, out, out_lp, lp_kind, B, nfrm,
                                                                       nsrl, nppf2, dv, dl);
    return check_launch("build_xmul");
}

// one warp per token row m = ((b*nfrm + f)*nsrl + s)*nppf2 + p'
__global__ void __launch_bounds__(256)
lin2_tail_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ w2, const float* __restrict__ b2,
                 const long long* __restrict__ srl_msk, const long long* __restrict__ cmp_msk,
                 float* __restrict__ logits, float* __restrict__ scores, int B, int nfrm, int nsrl, int nppf2,
                 int K, int ncmp, int nppf, int nfrm0, int spat)
{
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const long long M = (long long)B * nfrm * nsrl * nppf2;
    pdl_trigger();
    pdl_wait();
    if (row >= M) return;
    const float* hr = h + (size_t)row * ldh;
    float acc = 0.f;
    for (int c = lane * 4; c < K; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(hr + c);
        const float4 w = __ldg(reinterpret_cast<const float4*>(w2 + c));
        acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc); acc = fmaf(a.z, w.z, acc); acc = fmaf(a.w, w.w, acc);
    }
    acc = warp_sum(acc);
    if (lane != 0) return;
    long long m = row;
    const int pp = (int)(m % nppf2); m /= nppf2;
    const int s = (int)(m % nsrl); m /= nsrl;
    const int f = (int)(m % nfrm);
    const int b = (int)(m / nfrm);
    const int P = nfrm * nppf2;
    const int pidx = f * nppf2 + pp;                       // proposal index inside the query
    // which concatenated video the proposal belongs to (rows are [frame][vid][prop] for spat,
    // [vid][frame][prop] for temp)
    const int vid = spat ? (pidx / nppf) % ncmp : pidx / (nfrm0 * nppf);
    const float logit = acc + b2[0];
    const size_t o = ((size_t)b * nsrl + s) * P + pidx;
    logits[o] = logit;
    const float mk = (float)srl_msk[(size_t)b * nsrl + s] * (float)cmp_msk[(size_t)b * ncmp + vid];
    scores[o] = (1.f / (1.f + expf(-logit))) * mk;
}

int lin2_tail(const float* h, int ldh, const float* w2, const float* b2, const long long* srl_msk,
              const long long* cmp_msk, float* logits, float* scores, int B, int nfrm, int nsrl, int nppf2,
              int K, int ncmp, int nppf, int nfrm0, int spat, cudaStream_t st)
{
    VOG_REQUIRE(K % 4 == 0 && ldh % 4 == 0, "lin2_tail: K and ldh must be multiples of 4");
    const long long M = (long long)B * nfrm * nsrl * nppf2;
    if (M == 0) return 0;
    VOG_CUDA(launch_pdl(lin2_tail_kernel, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, st, h, ldh, w2, b2, srl_msk, cmp_msk,
                        logits, scores, B, nfrm, nsrl, nppf2, K, ncmp, nppf, nfrm0, spat));
    return check_launch("lin2_tail");
}

// -------------------------------------------------------------------------------------------------
// Language-side glue (code/mdl_vog.py:67-140): three kernels instead of ~15 library launches.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_lp4(void* base, size_t idx, float4 v, int kind) {
    if (kind == 0) {                                    // plain fp32 copy
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
    } else if (kind == 1) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) =
            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) =
            make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
}

// x_lp[t*Bq + b, :] = lp(emb[tok]),  tok = mask[b,t] == -1 ? pad_idx : words[b, mask[b,t]]
// (get_srl_arg_seq_to_sent_seq :67-95 + embed_tokens, time-major rows for the LSTM)
__global__ void __launch_bounds__(128)
lang_embed_kernel(const long long* __restrict__ words, int nwords, const long long* __restrict__ mask, int T,
                  const float* __restrict__ emb, int E, long long pad_idx, int Bq, void* __restrict__ out_lp,
                  int lp_kind)
{
    const int row = blockIdx.x;                 // t*Bq + b
    const int t = row / Bq, b = row % Bq;
    pdl_trigger();
    pdl_wait();
    const long long mk = mask[(size_t)b * T + t];
    long long tok = pad_idx;
    if (mk != -1) tok = words[(size_t)b * nwords + (mk < 0 ? 0 : (mk >= nwords ? nwords - 1 : mk))];
    // nn.Embedding raises on ids outside [0, pad_idx]; a kernel cannot, so out-of-range ids read the (zero) padding row
    if (tok < 0 || tok > pad_idx) tok = pad_idx;
    const float4* src = reinterpret_cast<const float4*>(emb + (size_t)tok * E);
    // blockIdx.y owns a 512-float4 column chunk: four independent 16-byte loads per thread, and enough CTAs to fill
    // the GPU when the rows are the 32 KB lines of the layer-0 projection table (80 rows at spat/gt5)
    const int c0 = blockIdx.y * 512 + threadIdx.x, n4 = E / 4;
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (c0 + 128 * k < n4) v[k] = __ldg(src + c0 + 128 * k);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (c0 + 128 * k < n4) store_lp4(out_lp, (size_t)row * E + 4 * (c0 + 128 * k), v[k], lp_kind);
}

int lang_embed(const long long* words, int nwords, const long long* mask, int T, const float* emb, int E,
               long long pad_idx, int Bq, void* out_lp, int lp_kind, cudaStream_t st)
{
    VOG_REQUIRE(E % 4 == 0, "lang_embed: embedding width must be a multiple of 4");
    if (T * Bq == 0) return 0;
    VOG_CUDA(launch_pdl(lang_embed_kernel, dim3(T * Bq, (E / 4 + 511) / 512), dim3(128), 0, st, words, nwords, mask, T, emb, E,
                        pad_idx, Bq, out_lp, lp_kind));
    return check_launch("lang_embed");
}

// out_lp[b*nsrl + s, :] = lp([full[cap[b,s,0]*Bq + b, :] | full[cap[b,s,1]*Bq + b, :]])
// (retrieve_srl_arg_from_lang_encode :97-131: first / last word of every SRL argument)
__global__ void __launch_bounds__(128)
lang_gather_kernel(const float* __restrict__ full, int D, const long long* __restrict__ cap, int T, int Bq,
                   int nsrl, void* __restrict__ out_lp, int lp_kind)
{
    const int row = blockIdx.x;                 // b*nsrl + s
    const int b = row / nsrl;
    pdl_trigger();
    pdl_wait();
    for (int c = threadIdx.x; c < 2 * D / 4; c += blockDim.x) {
        const int which = (4 * c) / D, col = 4 * c - which * D;
        long long t = cap[(size_t)row * 2 + which];
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        const float4 v = *reinterpret_cast<const float4*>(full + ((size_t)t * Bq + b) * D + col);
        store_lp4(out_lp, (size_t)row * 2 * D + 4 * c, v, lp_kind);
    }
}

int lang_gather(const float* full, int D, const long long* cap, int T, int Bq, int nsrl, void* out_lp, int lp_kind,
                cudaStream_t st)
{
    VOG_REQUIRE(D % 4 == 0, "lang_gather: feature width must be a multiple of 4");
    if (Bq * nsrl == 0) return 0;
    VOG_CUDA(launch_pdl(lang_gather_kernel, dim3(Bq * nsrl), dim3(128), 0, st, full, D, cap, T, Bq, nsrl, out_lp, lp_kind));
    return check_launch("lang_gather");
}

// out[r, :] = x[r, :] * (float)msk[r]  (+ low-precision copy)    (srl_arg_inds_msk :137-140)
__global__ void __launch_bounds__(128)
mask_rows_kernel(const float* __restrict__ x, const long long* __restrict__ msk, int D, float* __restrict__ out,
                 void* __restrict__ out_lp, int lp_kind)
{
    const int row = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    const float m = (float)msk[row];
    for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
        float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * D + 4 * c);
        v.x *= m; v.y *= m; v.z *= m; v.w *= m;
        *reinterpret_cast<float4*>(out + (size_t)row * D + 4 * c) = v;
        if (out_lp) store_lp4(out_lp, (size_t)row * D + 4 * c, v, lp_kind);
    }
}

int mask_rows(const float* x, const long long* msk, int rows, int D, float* out, void* out_lp, int lp_kind,
              cudaStream_t st)
{
    VOG_REQUIRE(D % 4 == 0, "mask_rows: feature width must be a multiple of 4");
    if (rows == 0) return 0;
    VOG_CUDA(launch_pdl(mask_rows_kernel, dim3(rows), dim3(128), 0, st, x, msk, D, out, out_lp, lp_kind));
    return check_launch("mask_rows");
}


// =============================================================================================
// Weight packing of one attention block for the tensor-core path: per-head zero-padded Wq | Wk | Wv rows
// ([3*H*dhp, d]: row (which*H + h)*dhp + r = W_which[off_h + r, :] for r < dh_h, else 0) and Wo with matching
// zero-padded columns ([d, H*dhp]), cast to bf16 or tf32-rounded fp32.  torch.chunk head split of
// code/transformer_code.py:169-186.
// =============================================================================================
struct PackParams { int d, H, dhp, kind; int off[VOG_MAX_HEADS], dh[VOG_MAX_HEADS]; };

__device__ __forceinline__ void pack_store(void* dst, long long idx, float v, int kind) {
    if (kind == 1) reinterpret_cast<__nv_bfloat16*>(dst)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(dst)[idx] = to_tf32(v);
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ wq, const float* __restrict__ wk, const float* __restrict__ wv,
                    const float* __restrict__ wo, void* __restrict__ wqkv, void* __restrict__ wo_p, PackParams p)
{
    const long long n1 = 3LL * p.H * p.dhp * p.d, n2 = (long long)p.d * p.H * p.dhp;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n1) {
        const int c = (int)(idx % p.d);
        const int row = (int)(idx / p.d);
        const int r = row % p.dhp, wh = row / p.dhp;
        const int h = wh % p.H, which = wh / p.H;
        const float* w = which == 0 ? wq : (which == 1 ? wk : wv);
        pack_store(wqkv, idx, r < p.dh[h] ? w[(size_t)(p.off[h] + r) * p.d + c] : 0.f, p.kind);
    } else if (idx < n1 + n2) {
        const long long j = idx - n1;
        const int col = (int)(j % (p.H * p.dhp));
        const int row = (int)(j / (p.H * p.dhp));
        const int h = col / p.dhp, r = col % p.dhp;
        pack_store(wo_p, j, r < p.dh[h] ? wo[(size_t)row * p.d + p.off[h] + r] : 0.f, p.kind);
    }
}

int pack_weights(const float* wq, const float* wk, const float* wv, const float* wo, int d, int H, const int* dh, int dhp,
                 int lp_kind, void* wqkv, void* wo_p, cudaStream_t st)
{
    VOG_REQUIRE(H >= 1 && H <= VOG_MAX_HEADS && d > 0 && dhp > 0, "pack_weights: bad geometry");
    VOG_REQUIRE(lp_kind == 1 || lp_kind == 2, "pack_weights: lp_kind must be VOG_LP_BF16 or VOG_LP_TF32");
    PackParams p;
    p.d = d; p.H = H; p.dhp = dhp; p.kind = lp_kind;
    int off = 0;
    for (int h = 0; h < H; ++h) {
        VOG_REQUIRE(dh[h] >= 1 && dh[h] <= dhp, "pack_weights: head dim %d does not fit dhp=%d", dh[h], dhp);
        p.off[h] = off; p.dh[h] = dh[h]; off += dh[h];
    }
    VOG_REQUIRE(off == d, "pack_weights: head dims sum to %d, d_model is %d", off, d);
    const long long n = 4LL * H * dhp * d;
    pack_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(wq, wk, wv, wo, wqkv, wo_p, p);
    return check_launch("pack_weights");
}

}  // namespace vog
