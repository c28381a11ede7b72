#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_loss.py -m gpu -x -q > gpurun_out/pytest_loss.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_loss.log
tail -25 gpurun_out/pytest_loss.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import vognet_pytorch_b200 as vb
from vognet_pytorch_b200 import synth
for name in ('spat_gt5', 'spat_p100'):
    w, batch = synth.workload(name)
    inp = dict(batch); inp.update(synth.make_loss_inputs(batch, **w))
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    d = {k: v.cuda() for k, v in inp.items()}
    lg = torch.randn(w['B'], 1, 5, d['pad_proposals'].shape[1], device='cuda')
    with torch.no_grad():
        for _ in range(3): fn({'mdl_outs': lg}, d)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn({'mdl_outs': lg}, d)
        e1.record(); torch.cuda.synchronize()
    print(name, 'loss forward', round(e0.elapsed_time(e1) / 20 * 1e3, 1), 'us per call')
PY
