#!/usr/bin/env python
"""HBM roofline of the fused Adam step at the model's size (45.25 M fp32 parameters): 16 B read + 12 B written per
element / CUDA-event time, L2 flushed between launches, against the measured copy bandwidth."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vognet_pytorch_b200.optim import FlatAdam  # noqa: E402

try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
p = torch.randn(45250728, device=dev).requires_grad_(True)
opt = FlatAdam([p], lr=1e-4, betas=(0.9, 0.99))
opt.flat_grad.normal_()
for _ in range(3):
    opt.step()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); opt.step(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
nbytes = 28 * opt.numel
print(f'vog_adam_step n={opt.numel}: {nbytes / 1e6:.1f} MB  {ts[5]:.1f} us  {nbytes / ts[5] / 1e3:.1f} GB/s = '
      f'{nbytes / ts[5] / 1e3 / peak:.3f} of the measured {peak:.0f} GB/s copy bandwidth')
ref = [torch.randn(s, device=dev, requires_grad=True) for s in [(1024, 1024)] * 44]
for r in ref:
    r.grad = torch.randn_like(r)
topt = torch.optim.Adam(ref, lr=1e-4, betas=(0.9, 0.99))
for _ in range(3):
    topt.step()
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); topt.step(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
print(f'torch.optim.Adam (default foreach path) on 44 x 1 Mi parameters ({44 * 1048576} elements): {ts[5]:.1f} us')
