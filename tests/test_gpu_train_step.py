"""FusedTrainStep (forward + loss + backward into the flat gradient + all-reduce + fused Adam) against the reference
trainer's protocol ``loss.backward(); optimizer.step()`` (utils/trn_utils.py:497-505, code/main_dist.py:55,75-80) on the
same kernels, and - on a box with >= 2 GPUs - the NCCL collectives of the path (gradient all-reduce, prediction gather
that replaces code/eval_vsrl_corr.py:125-140)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = 'cuda:0'


def _setup(compute, dev=DEV, seed=1):
    import vognet_pytorch_b200 as vb
    from vognet_pytorch_b200 import synth
    from vognet_pytorch_b200.optim import FlatAdam
    w, batch = synth.workload('cpu_ref', seed=seed)
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    sel = vb.get_mdl_loss_eval(cfg)
    mdl = sel['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    mdl = mdl.to(dev).set_compute(compute).train()
    mdl.train_dropout = False
    opt = FlatAdam(mdl.parameters(), lr=1e-3, betas=(0.9, 0.99))
    return mdl, sel['loss'](cfg, comm), opt, {k: v.to(dev) for k, v in inp.items()}


@pytest.mark.parametrize('compute', ['fp32x', 'bf16'])
def test_fused_step_equals_backward_plus_optimizer_step(compute):
    from vognet_pytorch_b200.train_step import FusedTrainStep
    mdl_a, loss_a, opt_a, inp = _setup(compute)
    mdl_b, loss_b, opt_b, _ = _setup(compute)
    step = FusedTrainStep(mdl_a, loss_a, opt_a)
    for it in range(2):
        la = step(inp)
        opt_b.zero_grad()
        lb = loss_b(mdl_b(inp), inp)['loss']
        lb.backward()
        # same kernels; the only differences are the order of the atomics inside the weight-gradient kernels and
        # one extra fp32 add per element (accumulate-into-zero vs assignment)
        ga, gb = opt_a.flat_grad, opt_b.flat_grad
        assert float((ga - gb).abs().max()) <= (1e-5 if compute == 'fp32x' else 1e-3) * float(gb.abs().max()), it
        opt_b.step()
        assert float(la) == pytest.approx(float(lb.detach()), rel=1e-5)
        # Adam normalises every coordinate: a gradient that is ~0 may flip sign between the two runs, so parameters
        # agree to a fraction of one update (lr = 1e-3), not to rounding
        for (k, pa), (_, pb) in zip(mdl_a.named_parameters(), mdl_b.named_parameters()):
            assert float((pa - pb).abs().max()) <= 2.1e-3 * (it + 1), k
    assert opt_a.step_count == opt_b.step_count == 2


def test_training_reduces_the_loss_on_a_fixed_batch():
    from vognet_pytorch_b200.train_step import FusedTrainStep
    mdl, loss_fn, opt, inp = _setup('bf16')
    mdl.train_dropout = True
    torch.manual_seed(0)
    step = FusedTrainStep(mdl, loss_fn, opt)
    losses = [float(step(inp)) for _ in range(12)]
    assert all(l == l for l in losses) and min(losses[-3:]) < 0.7 * losses[0], losses


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
import importlib.util
_spec = importlib.util.spec_from_file_location('vog_train_step_tests', os.path.join(sys.argv[1], 'tests', 'test_gpu_train_step.py'))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
_setup = _mod._setup
from vognet_pytorch_b200.train_step import FusedTrainStep
from vognet_pytorch_b200 import runtime
dev = f'cuda:{rank}'
mdl, loss_fn, opt, inp = _setup('bf16', dev, seed=1 + rank)
step = FusedTrainStep(mdl, loss_fn, opt)
# reference: gradients of both shards computed locally on every rank (deterministic: dropout off), averaged
grads = []
for r in range(world):
    m2, l2, o2, i2 = _setup('bf16', dev, seed=1 + r)
    o2.zero_grad()
    l2(m2(i2), i2)['loss'].backward()
    grads.append(o2.flat_grad.clone())
mean = sum(grads) / world
loss = step(inp)
# after the step flat_grad holds the SUM over ranks (the 1/world is folded into Adam)
err = float((opt.flat_grad / world - mean).abs().max() / mean.abs().max())
assert err < 1e-4, err
# every rank ends with identical parameters
chk = opt.flat_param.double().sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
assert all(float(c) == float(allc[0]) for c in allc), allc
# prediction gather over NCCL, ragged query counts
pred = {'boxes': torch.full((2 + rank, 3), float(rank), device=dev), 'idx': torch.arange(2 + rank, device=dev) + 10 * rank}
out = runtime.gather_predictions(pred)
assert out['boxes'].shape[0] == sum(2 + r for r in range(world))
assert out['idx'].tolist() == [i + 10 * r for r in range(world) for i in range(2 + r)]
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok', err)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs (NCCL)')
def test_nccl_gradient_allreduce_and_prediction_gather_two_ranks(tmp_path):
    import subprocess
    script = tmp_path / 'worker.py'
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', str(port), str(script), ROOT],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(' ok ') == 2
