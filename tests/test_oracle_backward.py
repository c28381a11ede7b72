"""Backward pin for SURVEY.md section 8f row 2: autograd through the ORACLE restatement (forward + loss) against the
parameter gradients torch autograd produced for the UNMODIFIED reference (tests/golden/grad_cpu_ref.npz, made by
oracle/make_golden.py: eval mode, LossB_SPAT, BASELINE config 1).  The CUDA backward of the model is not built yet;
this is the checker it will be held to, and it confirms that the restated forward is the same FUNCTION as the
reference (same value AND same derivative with respect to all 57 live parameters)."""
import os

import numpy as np

from oracle import vog_oracle as vo
from vognet_pytorch_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_oracle_parameter_gradients_match_reference():
    g = np.load(os.path.join(GOLD, 'grad_cpu_ref.npz'))
    w, batch = synth.workload('cpu_ref')
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    sd = {k: v.clone().requires_grad_(True) for k, v in synth.make_state_dict().items()}
    out = vo.vog_forward(sd, synth.clone_batch(inp), w['conc_type'], w['nppf'])
    loss = vo.loss_forward(out['mdl_outs'], inp, w['conc_type'], w['ncmp'], w['nppf'])['loss']
    assert abs(float(loss.detach()) - float(g['loss'])) <= 2e-6 * abs(float(g['loss']))
    loss.backward()
    unused = set(g['unused'].tolist())
    assert unused == {k for k, v in sd.items() if v.grad is None}      # the three heads the forward never reads
    checked = 0
    for k, v in sd.items():
        if k in unused:
            continue
        gr = v.grad.double().reshape(-1)
        norm = float(g['norm/' + k])
        assert abs(float(gr.norm()) - norm) <= 2e-4 * norm + 1e-9, k
        assert np.abs(gr[:8].numpy() - g['head/' + k]).max() <= 2e-4 * max(np.abs(g['head/' + k]).max(), norm * 1e-3), k
        assert abs(float(gr.sum()) - float(g['sum/' + k])) <= 1e-3 * norm * max(1.0, gr.numel() ** 0.5 * 1e-2) + 1e-9, k
        checked += 1
    assert checked == 57
