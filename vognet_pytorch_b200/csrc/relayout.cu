// Device-side contrastive-sample concatenation: the reference builds the SPAT / TEMP
// "one video with ncmp screens / ncmp clips" batch on the host, per sample, inside its dataset
// (code/dat_loader_simple.py:1067-1103,1147-1153,1196-1207 SPAT; :1231-1252,1290-1292 TEMP).  Here the per-video
// tensors [B,ncmp,...] (the SEP layout, the natural output of a loader) are uploaded once and concatenated on the GPU:
//
//   SPAT  rows [vid][frame][prop] -> [frame][vid][prop] (reshuffle_boxes) for proposals and region features,
//         segment features [vid][frame] -> [frame][vid]; proposal columns 0 and 2 (x1, x2) += 720 * vid
//   TEMP  row order unchanged (the [B,ncmp,P1,..] tensors ARE the concatenated ones); proposal column 4 (frame id)
//         += 10 * vid
//
// Pure HBM-bound byte movement: one CTA per destination row (8 KB / 12 KB), reads and writes fully coalesced, several
// 16-byte loads in flight per thread; the 28-byte proposal rows go element-wise.  The shifts are single fp32 adds of an exactly representable
// integer, the same operation `props + delta` performs in the reference - results are bit-identical.
#include "common.cuh"
#include "kernels.h"

namespace vog {

// dst row (b, f, v, p) <- src row (b, v, f, p); rows of `row4` float4 chunks.  per = rows per (video, frame) slot.
// One CTA per destination row: the index arithmetic happens once per CTA, every thread keeps up to four independent
// 16-byte loads in flight.
__global__ void __launch_bounds__(256)
permute_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int row4, int ncmp, int nfrm, int per)
{
    const long long row = blockIdx.x;
    const int p = (int)(row % per);
    long long r = row / per;
    const int v = (int)(r % ncmp); r /= ncmp;
    const int f = (int)(r % nfrm);
    const long long b = r / nfrm;
    const float4* s = src + (((b * ncmp + v) * nfrm + f) * per + p) * row4;
    float4* d = dst + row * row4;
    int c = threadIdx.x;
    for (; c + 768 < row4; c += 1024) {
        const float4 a0 = __ldg(s + c), a1 = __ldg(s + c + 256), a2 = __ldg(s + c + 512), a3 = __ldg(s + c + 768);
        d[c] = a0; d[c + 256] = a1; d[c + 512] = a2; d[c + 768] = a3;
    }
    for (; c + 256 < row4; c += 512) {
        const float4 a0 = __ldg(s + c), a1 = __ldg(s + c + 256);
        d[c] = a0; d[c + 256] = a1;
    }
    for (; c < row4; c += 256) d[c] = __ldg(s + c);
}

// proposals: pdim floats per row (7: not 16-byte aligned), one thread per element
__global__ void __launch_bounds__(256)
concat_props_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total, int pdim, int ncmp,
                    int nfrm, int nppf, int spat, float shift)
{
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const long long row = i / pdim;
    const int c = (int)(i - row * pdim);
    const int p = (int)(row % nppf);
    long long r = row / nppf;
    int v, f;
    if (spat) { v = (int)(r % ncmp); r /= ncmp; f = (int)(r % nfrm); r /= nfrm; }
    else      { f = (int)(r % nfrm); r /= nfrm; v = (int)(r % ncmp); r /= ncmp; }
    const long long srow = ((r * ncmp + v) * nfrm + f) * nppf + p;
    float x = src[srow * pdim + c];
    const bool shifted = spat ? (c == 0 || c == 2) : (c == 4);
    if (shifted) x = __fadd_rn(x, __fmul_rn((float)v, shift));        // delta = arange(n) * shift, exact in fp32
    dst[i] = x;
}

int concat_videos(const float* feat, int D, const float* seg, int Ds, const float* props, int pdim, float* feat_out,
                  float* seg_out, float* props_out, int B, int ncmp, int nfrm, int nppf, int spat, float shift,
                  cudaStream_t st)
{
    VOG_REQUIRE(B >= 0 && ncmp >= 1 && nfrm >= 1 && nppf >= 1, "concat_videos: bad dimension");
    if (B == 0) return 0;
    if (feat_out) {
        VOG_REQUIRE(feat && D > 0 && D % 4 == 0, "concat_videos: region features need D %% 4 == 0");
        VOG_REQUIRE(spat, "concat_videos: TEMP keeps the row order - use the input tensor as [B, ncmp*P1, D]");
        const long long rows = (long long)B * ncmp * nfrm * nppf;
        VOG_REQUIRE(rows < (1LL << 31), "concat_videos: too many proposal rows");
        permute_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(
            reinterpret_cast<const float4*>(feat), reinterpret_cast<float4*>(feat_out), D / 4, ncmp, nfrm, nppf);
        if (check_launch("concat_videos(feat)")) return -1;
    }
    if (seg_out) {
        VOG_REQUIRE(seg && Ds > 0 && Ds % 4 == 0, "concat_videos: segment features need Ds %% 4 == 0");
        VOG_REQUIRE(spat, "concat_videos: TEMP keeps the segment order - use the input tensor as [B, ncmp*nfrm, Ds]");
        permute_rows_kernel<<<(unsigned)(B * ncmp * nfrm), 256, 0, st>>>(
            reinterpret_cast<const float4*>(seg), reinterpret_cast<float4*>(seg_out), Ds / 4, ncmp, nfrm, 1);
        if (check_launch("concat_videos(seg)")) return -1;
    }
    if (props_out) {
        VOG_REQUIRE(props && pdim >= 5, "concat_videos: proposals need >= 5 columns");
        const long long t = (long long)B * ncmp * nfrm * nppf * pdim;
        concat_props_kernel<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(props, props_out, t, pdim, ncmp, nfrm, nppf,
                                                                        spat, shift);
        if (check_launch("concat_videos(props)")) return -1;
    }
    return 0;
}

}  // namespace vog
