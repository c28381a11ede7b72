#!/usr/bin/env python
"""Host-side cost per step of the public forward (graph replay) and of the e2e loop pieces: the GPU is idle most of
the time at the gt5 sizes, so what the host spends per call bounds the end-to-end rate."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vognet_pytorch_b200 as vb
from vognet_pytorch_b200 import synth

dev = 'cuda:0'
w, batch = synth.workload(sys.argv[1] if len(sys.argv) > 1 else 'spat_gt5')
cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
sel = vb.get_mdl_loss_eval(cfg)
m = sel['mdl'](cfg, comm)
m.load_state_dict(synth.make_state_dict(), strict=True)
m = m.to(dev).eval().set_compute('tf32')
m.use_cuda_graph = True
ev = sel['eval'](cfg, comm, dev)
res = {k: v.to(dev) for k, v in batch.items()}
with torch.no_grad():
    res = m.graph_input_buffers(res)
    for _ in range(5):
        out = m(res); s = ev.get_out_results_boxes(out, res)
torch.cuda.synchronize()


def timeit(fn, n=300):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6


with torch.no_grad():
    print('forward (graph replay)          host %.1f us/call, wall %.1f us/call' % timeit(lambda: m(res)))
    print('forward + selection             host %.1f us/call, wall %.1f us/call' % timeit(lambda: ev.get_out_results_boxes(m(res), res)))
    print('weights signature               host %.1f us/call' % timeit(lambda: m._weights_sig())[0])
    g = m._graph_for(res, res['new_srl_idxs'].shape[1])
    print('graph.replay() alone            host %.1f us/call, wall %.1f us/call' % timeit(lambda: g['graph'].replay()))
    print('clone of the outputs            host %.1f us/call' % timeit(lambda: {k: v.clone() for k, v in g['out'].items()})[0])
