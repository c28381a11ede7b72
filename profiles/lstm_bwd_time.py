#!/usr/bin/env python
"""LSTM backward through time, one layer (T=20, Bq=4, H=1024): persistent weight-resident kernel vs T per-step launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops, ops_bwd as ob, _lib  # noqa: E402

dev = 'cuda:0'
L = _lib.lib()
T, H = 20, 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Bq, lens in ((4, [7, 18, 11, 7]), (4, [20, 20, 20, 20]), (1, [18])):
    g = torch.Generator().manual_seed(0)
    gx = (torch.randn(T * Bq, 8 * H, generator=g) * 0.5).to(dev)
    whh = (torch.randn(2, 4 * H, H, generator=g) / 48).to(dev)
    lens_d = torch.tensor(lens, device=dev)
    _, acts = ops.lstm_layer_fwd(gx, whh, lens_d, T, Bq, ops.LP_NONE, want_acts=True)
    dout = torch.randn(T * Bq, 2 * H, generator=g).to(dev)
    wt = whh.transpose(1, 2).contiguous()
    for name, on in (('per-step launches', 0), ('persistent kernel', 1)):
        L.vog_debug_lstm_bwd_resident(on)
        for _ in range(3):
            ob.lstm_bwd_steps(dout, acts, wt, lens_d, T, Bq, whh=whh)
        ts = []
        for cold in (False, True):
            ts = []
            for _ in range(10):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ob.lstm_bwd_steps(dout, acts, wt, lens_d, T, Bq, whh=whh)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            print(f'Bq={Bq} lens={lens} {name}: median {sorted(ts)[5]:.1f} us ({"L2 flushed" if cold else "warm"})')
    L.vog_debug_lstm_bwd_resident(1)
    import ctypes
    buf = torch.zeros(8, dtype=torch.int64, device=dev)
    L.vog_debug_lstm_trace(ctypes.c_void_p(buf.data_ptr()))
    ob.lstm_bwd_steps(dout, acts, wt, lens_d, T, Bq, whh=whh)
    torch.cuda.synchronize()
    L.vog_debug_lstm_trace(None)
    v = buf.cpu().tolist()
    if v[5]:
        n = v[5]
        print(f'   traced launch (CTA 0 / thread 0), {n} steps; per step (cycles): collect {v[0] / n:.0f}  gate grads {v[1] / n:.0f}  '
              f'barrier {v[2] / n:.0f}  matvec {v[3] / n:.0f}  publish {v[4] / n:.0f}')
