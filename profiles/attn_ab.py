#!/usr/bin/env python
"""A/B of the two fused-attention kernels (v1: Q/P through shared memory, v2: Q/P in tensor memory) at the
spat/p100 and gt5 shapes: CUDA-event time with L2 flushed, max |difference| between the two."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops, _lib  # noqa: E402

dev = 'cuda:0'
L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Bt, N, d in ((40, 2000, 768), (4, 4000, 512), (40, 100, 768), (4, 200, 512), (400, 500, 768)):
    hd = ops.chunk_sizes(d, 3)
    dhp = ops.round_up(max(hd), 64)
    q = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    k = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    v = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    nbox = N // 5 if N % 5 == 0 else N
    a = torch.rand(Bt * nbox, 3, device=dev)
    bpe = torch.zeros(3, device=dev)
    outs = {}
    qt = -(-N // 128)
    for impl, cl in ((1, 0), (2, 1), (3, 1)):
        if cl and qt % cl:
            continue
        L.vog_debug_attn_impl(impl)
        L.vog_debug_attn_cluster(cl)
        out = torch.empty(Bt * N, 3 * dhp, device=dev, dtype=torch.bfloat16)
        run = lambda: ops.tc_attn_fwd(q, k, v, N, hd, 1.0 / d ** 0.5, out=out, bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe)
        for _ in range(3):
            run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        if impl == 1:
            outs[1] = out.float().clone()
        else:
            assert (outs[1] - out.float()).abs().max().item() < 1e-2, f'v{impl} differs from v1'
            outs[2] = out.float().clone()
        fl = 4.0 * Bt * N * N * d
        print(f'Bt={Bt:4d} N={N:5d} d={d} impl v{impl} cluster {cl}: {ts[5]:8.1f} us  {fl / ts[5] / 1e6:7.1f} TF/s')
    print(f'    max |v2 - v1| = {(outs[1] - outs[2]).abs().max().item():.3e}')
L.vog_debug_attn_impl(2)
L.vog_debug_attn_cluster(0)
