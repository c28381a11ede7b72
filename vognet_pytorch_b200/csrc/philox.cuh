// Counter-based random stream of the training-mode dropout masks (Philox4x32, 7 rounds - the shortest variant that
// passes BigCrush, Salmon et al. SC'11).  A mask bit is a pure function of (seed, stream id, row, column), so the
// backward kernels regenerate exactly the mask the forward drew and nothing N x N is stored:
//   attention probabilities   code/transformer_code.py:153   (stream = sequence*H + head, row = query, column = key)
//   residual branches         code/transformer_code.py:26,31 (stream = call-site id, row = token row, column = feature)
//   LSTM input / between layers / output   utils/mdl_srl_utils.py:104,128,150
// One Philox call yields 128 bits = eight 16-bit uniforms: element e of the aligned group of eight columns is kept iff
// its 16-bit uniform >= round(p * 65536)  (p is realised to within 2^-17; torch's own stream cannot be matched bit for
// bit anyway - parity of the dropout path is statistical + forward/backward mask consistency).
#pragma once
#include <stdint.h>

namespace vog {

__host__ __device__ inline uint32_t drop_threshold16(float p) {
    const float t = p * 65536.f + 0.5f;
    return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : (uint32_t)t);
}

__device__ __forceinline__ void philox4x32_7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                             uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// eight 16-bit uniforms for columns col0 .. col0+7 (col0 % 8 == 0) of (stream, row): out[e >> 1] >> 16*(e & 1)
__device__ __forceinline__ void attn_rand16x8(unsigned long long seed, uint32_t stream, uint32_t row, uint32_t col0,
                                              uint32_t* out)
{
    uint32_t r[4];
    philox4x32_7(col0 >> 3, row, stream, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; out[3] = r[3];
}

// Attention probabilities (N x N elements per head: the only call site where the generator itself shows up in the
// profile) use 8-bit uniforms, sixteen per Philox call: keep iff byte >= round(p * 256), and 1 / keep-probability is
// taken from the REALISED threshold so that E[mask / (1 - p_eff)] = 1 exactly (p = 0.2 -> p_eff = 51/256 = 0.1992).
__host__ __device__ inline uint32_t drop_threshold8(float p) {
    const float t = p * 256.f + 0.5f;
    return t <= 0.f ? 0u : (t >= 255.f ? 255u : (uint32_t)t);
}
__host__ __device__ inline float drop_inv_keep8(float p) { return 256.f / (256.f - (float)drop_threshold8(p)); }

// sixteen 8-bit uniforms for columns col0 .. col0+15 (col0 % 16 == 0) of (stream, row): byte e & 3 of out[e >> 2]
__device__ __forceinline__ void attn_rand8x16(unsigned long long seed, uint32_t stream, uint32_t row, uint32_t col0,
                                              uint32_t (&out)[4])
{
    philox4x32_7(col0 >> 4, row, stream, 0xa77eu, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

}  // namespace vog
