#!/usr/bin/env python
"""clock64 phase breakdown of CTA 0 of the weight-resident LSTM recurrence kernel."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops, _lib  # noqa: E402

dev = 'cuda:0'
L = _lib.lib()
T, H = 20, 1024
NAMES = {0: 'tagged', 1: 'flags', 2: 'records', 3: 'records, two units per warp'}
CASES = [(4, 0), (4, 2), (4, 3), (8, 2), (1, 2), (1, 3), (2, 2), (2, 3), (3, 2), (3, 3)]
for Bq, mode in CASES:
    xmode, backoff = mode & 0xff, mode >> 8
    L.vog_debug_lstm_exchange(mode)
    gx = torch.rand(T * Bq, 8 * H, device=dev) - 0.5
    whh = (torch.rand(2, 4 * H, H, device=dev) - 0.5) / 32
    lens = torch.tensor([7, 18, 11, 7, 20, 3, 9, 14], device=dev)[:Bq]
    if Bq <= 2:
        lens = torch.tensor([18, 11], device=dev)[:Bq]
    buf = torch.zeros(8, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32)
    torch.cuda.synchronize()
    L.vog_debug_lstm_trace(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32)
    e1.record()
    torch.cuda.synchronize()
    L.vog_debug_lstm_trace(None)
    v = buf.cpu().tolist()
    n = max(v[5], 1)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_TF32)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    tag = f'Bq={Bq} exchange={NAMES[xmode]}' + (f' backoff/delay={backoff}' if backoff else '')
    print(f'{tag}: median {sorted(ts)[10]:.1f} us, min {min(ts):.1f} us (20 launches incl. the zero kernel)')
    if v[5]:
        print(f'   traced launch: {v[5]} steps; per step (cycles): '
              f'matvec {v[0] / n:.0f}  reduce {v[1] / n:.0f}  cell+publish {v[2] / n:.0f}  poll {v[3] / n:.0f}  barrier {v[4] / n:.0f}'
              f'   | entry -> first step {v[6]} cycles')
L.vog_debug_lstm_exchange(4)
