#!/usr/bin/env python
"""(packed batches) Where the end-to-end step goes: CUDA events around the pieces of one e2e step on the compute stream (wait for H2D,
staging + graph replay, selection, D2H) over 100 pipelined steps, plus the step period."""
import os, sys, time, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vognet_pytorch_b200 as vb
from vognet_pytorch_b200 import synth
from vognet_pytorch_b200.runtime import BatchPrefetcher, pack_host_batch

dev = torch.device('cuda', 0)
w, batch = synth.workload('spat_gt5')
cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
sel = vb.get_mdl_loss_eval(cfg)
m = sel['mdl'](cfg, comm)
m.load_state_dict(synth.make_state_dict(), strict=True)
m = m.to(dev).eval().set_compute('tf32')
m.use_cuda_graph = True
ev = sel['eval'](cfg, comm, dev)
host = {k: v.pin_memory() for k, v in batch.items()}
if 'plain' not in sys.argv:
    host = pack_host_batch(host, first=tuple(k for k in m._GRAPH_KEYS if k in host))
N = 120
pre = BatchPrefetcher((host for _ in range(N + 10)), dev)
host_out = [None, None]
done = [None, None]
marks = []


def E():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def launch(i):
    e0 = E()
    b = pre.next()
    e1 = E()
    out = m(b)
    e2 = E()
    s = ev.get_out_results_boxes(out, b)
    e3 = E()
    res = (s['boxes'], s['scores'], s['indexs'])
    if host_out[i & 1] is None:
        host_out[i & 1] = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res]
    for h_, r in zip(host_out[i & 1], res):
        h_.copy_(r, non_blocking=True)
    e4 = E()
    done[i & 1] = e4
    pre.release(b, e4)
    marks.append((e0, e1, e2, e3, e4))


with torch.no_grad():
    for i in range(10):
        launch(i)
        if i:
            done[(i - 1) & 1].synchronize()
    torch.cuda.synchronize()
    marks.clear()
    t0 = time.perf_counter()
    hs = []
    for i in range(N):
        th = time.perf_counter()
        launch(i)
        hs.append(time.perf_counter() - th)
        if i:
            done[(i - 1) & 1].synchronize()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / N
med = lambda xs: statistics.median(xs)
print(f'e2e wall per step {wall * 1e6:.1f} us  ({w["B"] / wall:.0f} queries/s); host time inside launch() median {med(hs) * 1e6:.1f} us')
print('GPU-side medians (us): wait-for-H2D %.1f | staging+graph+clone %.1f | selection %.1f | D2H enqueue %.1f' % tuple(
    med([a[k].elapsed_time(a[k + 1]) * 1e3 for a in marks[5:]]) for k in range(4)))
per = [marks[i][0].elapsed_time(marks[i + 1][0]) * 1e3 for i in range(5, N - 1)]
print('step period on the GPU timeline: median %.1f us, p90 %.1f us' % (med(per), sorted(per)[int(0.9 * len(per))]))
gap = [marks[i][4].elapsed_time(marks[i + 1][0]) * 1e3 for i in range(5, N - 1)]
print('idle gap between the end of step i and the first event of step i+1: median %.1f us' % med(gap))
