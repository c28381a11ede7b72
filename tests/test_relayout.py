"""Contrastive-sample concatenation on the device (vog_concat_videos).  CPU: the oracle
restatement against golden vectors produced by the reference's OWN host helpers (reshuffle_boxes / process_props lifted
from code/dat_loader_simple.py by oracle/ref_harness.py, see oracle/make_golden.py).  GPU: the kernel against the
oracle, bit-exact (pure data movement + one exact fp32 add), at the fixture sizes and at the BASELINE p100 size; and
the concatenated batch through the SPAT / TEMP models."""
import os

import numpy as np
import pytest
import torch

import vognet_pytorch_b200 as vb
from vognet_pytorch_b200 import synth
from oracle import vog_oracle as vo          # checker

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
VIS = ('pad_region_feature', 'seg_feature_for_frms', 'pad_proposals')


@pytest.mark.parametrize('name', list(synth.WORKLOADS_SEP))
@pytest.mark.parametrize('conc', ['spat', 'temp'])
def test_oracle_concat_matches_reference(name, conc):
    w, b = synth.workload(name)
    g = np.load(os.path.join(GOLD, f'relayout_{name}.npz'))
    f, s, p = vo.concat_videos(*(b[k] for k in VIS), conc, synth.NFRM0, w['nppf'])
    assert np.array_equal(p.numpy(), g[f'{conc}_props'])
    assert np.array_equal(f[..., :4].numpy(), g[f'{conc}_feat_head'])
    assert np.array_equal(f.double().sum(-1).numpy(), g[f'{conc}_feat_sum'])
    assert np.array_equal(s[..., :4].numpy(), g[f'{conc}_seg_head'])
    assert np.array_equal(s.double().sum(-1).numpy(), g[f'{conc}_seg_sum'])


def test_oracle_concat_layout_properties():
    """spat: row (f, v, p) of the output is row (v, f, p) of video v with x shifted by 720 v; temp: frame ids of
    video v live in [10 v, 10 v + 9]."""
    w, b = synth.workload('sep_gt5')
    nppf, ncmp = w['nppf'], w['ncmp']
    f, s, p = vo.concat_videos(*(b[k] for k in VIS), 'spat', 10, nppf)
    p5 = p.view(-1, 10, ncmp, nppf, 7)
    assert torch.equal(p5[:, 3, 2, 1, 1], b['pad_proposals'][:, 2, 3 * nppf + 1, 1])
    assert torch.equal(p5[:, 3, 2, 1, 0], b['pad_proposals'][:, 2, 3 * nppf + 1, 0] + 1440.0)
    assert torch.equal(s.view(-1, 10, ncmp, 3072)[:, 7, 1], b['seg_feature_for_frms'][:, 1, 7])
    _, _, pt = vo.concat_videos(*(b[k] for k in VIS), 'temp', 10, nppf)
    fr = pt.view(-1, ncmp, 10 * nppf, 7)[..., 4]
    for v in range(ncmp):
        assert fr[:, v].min() == 10 * v and fr[:, v].max() == 10 * v + 9


@pytest.mark.gpu
@pytest.mark.parametrize('case', [('sep_gt5', None), ('sep_p100', None), ('big', dict(B=4, ncmp=4, nppf=100))])
@pytest.mark.parametrize('conc', ['spat', 'temp'])
def test_cuda_concat_bit_exact(case, conc):
    name, kw = case
    if kw is None:
        w, b = synth.workload(name)
    else:                       # the spat/p100 BASELINE size: 4 queries x 4 videos x 1000 proposals
        w, b = kw, synth.make_batch_sep(seed=3, **kw)
    from vognet_pytorch_b200 import ops
    ref = vo.concat_videos(*(b[k] for k in VIS), conc, synth.NFRM0, w['nppf'])
    got = ops.concat_videos(*(b[k].cuda() for k in VIS), conc, synth.NFRM0, w['nppf'])
    torch.cuda.synchronize()
    for r, g_ in zip(ref, got):
        assert r.shape == g_.shape and torch.equal(r, g_.cpu())


@pytest.mark.gpu
@pytest.mark.parametrize('conc', ['spat', 'temp'])
def test_concat_batch_feeds_the_concatenated_models(conc):
    """A per-video batch uploaded once serves the SEP model directly and, through concat_batch, the SPAT / TEMP
    models: same result as running them on the host-concatenated batch (the oracle's concatenation)."""
    from vognet_pytorch_b200 import runtime
    w, b = synth.workload('sep_gt5')
    cfg, comm = synth.default_cfg(conc), synth.default_comm(w['nppf'])
    mdl = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    mdl = mdl.cuda().eval().set_compute('tf32')
    dev = {k: v.cuda() for k, v in b.items()}
    got = mdl(runtime.concat_batch(dev, conc, synth.NFRM0, w['nppf']))
    f, s, p = vo.concat_videos(*(b[k] for k in VIS), conc, synth.NFRM0, w['nppf'])
    host = {k: (v[:, :1].contiguous() if v.dim() > 1 and k.startswith('srl_') else v) for k, v in b.items()}
    host.update(pad_region_feature=f, seg_feature_for_frms=s, pad_proposals=p)
    host.pop('verb_ind_in_srl')
    ref = mdl({k: v.cuda() for k, v in host.items()})
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
    assert got['mdl_outs'].shape == (w['B'], 1, 5, w['ncmp'] * 10 * w['nppf'])


@pytest.mark.gpu
def test_concat_rejects_bad_shapes():
    from vognet_pytorch_b200 import ops
    w, b = synth.workload('sep_gt5')
    with pytest.raises(ValueError):
        ops.concat_videos(*(b[k].cuda() for k in VIS), 'spat', synth.NFRM0, w['nppf'] + 1)
    with pytest.raises(ValueError):
        ops.concat_videos(*(b[k].cuda() for k in VIS), 'sep', synth.NFRM0, w['nppf'])
