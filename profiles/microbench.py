#!/usr/bin/env python
"""Per-kernel CUDA-event timings of the hot-path ops at the BASELINE shapes (warm L2 and with a
256 MB flush between launches).  usage: python profiles/microbench.py [gt5|p100] [tf32|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops  # noqa: E402

dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20, cold=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm_case(name, M, N, K, kind, residual=False, bias=True, relu=False, f32=True, lp=False, rep=1):
    dt = torch.bfloat16 if kind == ops.LP_BF16 else torch.float32
    a = (torch.rand(M, K, device=dev) - 0.5).to(dt)
    w = (torch.rand(N, K, device=dev) - 0.5).to(dt)
    b = torch.rand(N, device=dev) if bias else None
    r = torch.rand(M, N, device=dev) if residual else None
    o32 = torch.empty(M * rep, N, device=dev) if f32 else None
    olp = torch.empty(M * rep, N, device=dev, dtype=dt) if lp else None

    def fn():
        ops.tc_gemm(a, w, bias=b, residual=r, relu=relu, out_f32=o32, out_lp=olp, rep=rep, want_f32=f32,
                    lp_kind=(kind if lp else ops.LP_NONE))
    tw, tc = timeit(fn), timeit(fn, cold=True)
    fl = 2.0 * M * N * K
    print(f'{name:28s} M={M:6d} N={N:5d} K={K:5d}  warm {tw:8.1f} us ({fl / tw / 1e6:7.1f} TF/s)  cold {tc:8.1f} us')


def attn_case(name, Bt, N, d, H=3):
    hd = ops.chunk_sizes(d, H)
    dhp = ops.round_up(max(hd), 64)
    q = (torch.rand(Bt, H, N, dhp, device=dev) - 0.5).bfloat16()
    k = (torch.rand(Bt, H, N, dhp, device=dev) - 0.5).bfloat16()
    vt = (torch.rand(Bt, H, N, dhp, device=dev) - 0.5).bfloat16()
    nbox = N // 5 if N % 5 == 0 else N
    a = torch.rand(Bt * nbox, H, device=dev)
    bpe = torch.zeros(H, device=dev)
    out = torch.empty(Bt * N, H * dhp, device=dev, dtype=torch.bfloat16)

    def fn():
        ops.tc_attn_fwd(q, k, vt, N, hd, 1.0 / d ** 0.5, out=out, bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe)
    tw, tc = timeit(fn), timeit(fn, cold=True)
    fl = 4.0 * Bt * N * N * d
    print(f'{name:28s} Bt={Bt:4d} N={N:5d} d={d}  warm {tw:8.1f} us ({fl / tw / 1e6:7.1f} TF/s)  cold {tc:8.1f} us')


def lstm_case(T=20, Bq=4, H=1024, kind=ops.LP_TF32):
    gx = torch.rand(T * Bq, 8 * H, device=dev) - 0.5
    whh = (torch.rand(2, 4 * H, H, device=dev) - 0.5) / 32
    lens = torch.tensor([7, 18, 11, 7], device=dev)[:Bq]

    def fn():
        ops.lstm_layer_fwd(gx, whh, lens, T, Bq, kind)
    print(f'lstm layer T={T} Bq={Bq} (max len 18)  warm {timeit(fn):8.1f} us  cold {timeit(fn, cold=True):8.1f} us')


def ln_case(M, d, kind):
    x = torch.rand(M, d, device=dev)
    w, b = torch.rand(d, device=dev), torch.rand(d, device=dev)
    o = torch.empty(M, d, device=dev)
    olp = torch.empty(M, d, device=dev, dtype=torch.bfloat16 if kind == ops.LP_BF16 else torch.float32)

    def fn():
        ops.add_layernorm(x, None, w, b, out=o, out_lp=olp, lp_kind=kind)
    t = timeit(fn, cold=True)
    by = M * d * (8 + (2 if kind == ops.LP_BF16 else 4))
    print(f'add_layernorm M={M} d={d}  cold {t:8.1f} us  ({by / t / 1e3:6.0f} GB/s)')


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'gt5'
    kind = ops.LP_BF16 if (len(sys.argv) > 2 and sys.argv[2] == 'bf16') else (ops.LP_TF32 if which == 'gt5' else ops.LP_BF16)
    B, nppf = 4, (5 if which == 'gt5' else 100)
    P = 40 * nppf
    Mo, Mm = B * P, B * 10 * 5 * 4 * nppf
    print(f'== {which} kind={"bf16" if kind == ops.LP_BF16 else "tf32"}  obj rows {Mo}  mul rows {Mm}')
    gemm_case('prop_encoder', Mo, 256, 2048, kind, relu=True, lp=True)
    gemm_case('seg_encoder (rep)', B * 40, 256, 3072, kind, relu=True, lp=True, rep=nppf)
    gemm_case('obj qkv (as plain, bf16 out)', Mo, 1728, 512, kind, bias=False, f32=False, lp=True)
    gemm_case('obj wo + residual', Mo, 512, 576, kind, bias=False, residual=True)
    gemm_case('obj ffn1', Mo, 256, 512, kind, relu=True, f32=False, lp=True)
    gemm_case('obj ffn2 + residual', Mo, 512, 256, kind, residual=True)
    gemm_case('mul qkv (as plain, bf16 out)', Mm, 2304, 768, kind, bias=False, f32=False, lp=True)
    gemm_case('mul wo + residual', Mm, 768, 768, kind, bias=False, residual=True)
    gemm_case('mul ffn1', Mm, 384, 768, kind, relu=True, f32=False, lp=True)
    gemm_case('mul ffn2 + residual', Mm, 768, 384, kind, residual=True)
    gemm_case('lin2.0', Mm, 256, 768, kind, relu=True)
    gemm_case('lstm gx l0', 80, 8192, 512, kind)
    gemm_case('lstm gx l1', 80, 8192, 2048, kind)
    attn_case('obj attention', B, P, 512)
    attn_case('mul attention', B * 10, 5 * 4 * nppf, 768)
    lstm_case(kind=kind)
    ln_case(Mo, 512, kind)
    ln_case(Mm, 768, kind)
