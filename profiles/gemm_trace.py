#!/usr/bin/env python
"""clock64 timeline of CTA 0 of the tcgen05 GEMM at small (latency-bound) shapes: where do the
microseconds of a tiny launch go?  usage: python profiles/gemm_trace.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops, _lib  # noqa: E402

dev = 'cuda:0'
L = _lib.lib()
NAMES = ['entry', 'setup done', 'first TMA issued', 'first stage landed', 'last MMA committed',
         'accumulator visible', 'epilogue done', 'exit']
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K, kind, cold) in ((20, 256, 512, ops.LP_TF32, 0), (20, 256, 512, ops.LP_TF32, 1), (800, 256, 512, ops.LP_TF32, 0),
                              (800, 256, 512, ops.LP_TF32, 1), (4000, 768, 768, ops.LP_TF32, 1), (4000, 2304, 768, ops.LP_BF16, 1),
                              (80, 8192, 2048, ops.LP_TF32, 1)):
    dt = torch.bfloat16 if kind == ops.LP_BF16 else torch.float32
    a = (torch.rand(M, K, device=dev) - 0.5).to(dt)
    w = (torch.rand(N, K, device=dev) - 0.5).to(dt)
    b = torch.rand(N, device=dev)
    o = torch.empty(M, N, device=dev)
    buf = torch.zeros(8, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.tc_gemm(a, w, bias=b, out_f32=o)
    torch.cuda.synchronize()
    L.vog_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
    if cold:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.tc_gemm(a, w, bias=b, out_f32=o)
    e1.record()
    torch.cuda.synchronize()
    L.vog_debug_gemm_trace(None)
    v = buf.cpu().tolist()
    print(f'M={M} N={N} K={K} {"bf16" if kind == ops.LP_BF16 else "tf32"} {"cold" if cold else "warm"}: '
          f'events {e0.elapsed_time(e1) * 1e3:.1f} us (incl. split-K reduce + host gaps), CTA0 lifetime {(v[7] - v[0]) / 1.9e3:.1f} us')
    for n, t in zip(NAMES, v):
        print(f'    {n:24s} +{(t - v[0]):8d} cycles')
