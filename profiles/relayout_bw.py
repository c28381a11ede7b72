#!/usr/bin/env python
"""HBM roofline of the device-side concatenation (vog_concat_videos) at the spat/p100 BASELINE size: algorithmic bytes
(every input byte read once, every output byte written once) / CUDA-event time, L2 flushed between launches, against
the measured copy bandwidth in MEASURED_PEAKS.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vognet_pytorch_b200 import ops  # noqa: E402

try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
dev = 'cuda:0'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B, ncmp, nppf in ((4, 4, 100), (32, 4, 100), (4, 4, 5)):
    P1 = 10 * nppf
    feat = torch.rand(B, ncmp, P1, 2048, device=dev)
    seg = torch.rand(B, ncmp, 10, 3072, device=dev)
    props = torch.rand(B, ncmp, P1, 7, device=dev)
    nbytes = 2 * 4 * (feat.numel() + seg.numel() + props.numel())
    for _ in range(3):
        ops.concat_videos(feat, seg, props, 'spat', 10, nppf)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.concat_videos(feat, seg, props, 'spat', 10, nppf); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    gbs = nbytes / ts[5] / 1e3
    print(f'concat_videos spat B={B} ncmp={ncmp} nppf={nppf}: {nbytes / 1e6:8.1f} MB  {ts[5]:8.1f} us  {gbs:7.1f} GB/s  '
          f'= {gbs / peak:.3f} of the measured {peak:.0f} GB/s copy bandwidth (3 launches + output allocation inside the timed region)')
