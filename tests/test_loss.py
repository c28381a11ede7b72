"""Grounding loss.  CPU: the oracle restatement against golden values of the
UNMODIFIED reference LossB_SPAT / LossB_TEMP (tests/golden/loss_*.npz, made by oracle/make_golden.py).  GPU:
vog_loss_fwd through the LossB_* modules against the same fixtures - boolean IoU targets bit-exact, loss value
within 2e-6 relative (transcendentals + summation order)."""
import os

import numpy as np
import pytest
import torch

import vognet_pytorch_b200 as vb
from vognet_pytorch_b200 import synth
from oracle import vog_oracle as vo          # checker

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = list(synth.WORKLOADS)


def _case(name):
    w, batch = synth.workload(name)
    g = np.load(os.path.join(GOLD, f'{name}.npz'))
    gl = np.load(os.path.join(GOLD, f'loss_{name}.npz'))
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    tg = np.unpackbits(gl['targets'])[:int(np.prod(gl['targets_shape']))].reshape(gl['targets_shape']).astype(bool)
    return w, inp, torch.from_numpy(g['mdl_outs']), float(gl['loss']), tg


@pytest.mark.parametrize('name', NAMES)
def test_oracle_loss_matches_reference(name):
    w, inp, logits, loss, tg = _case(name)
    r = vo.loss_forward(logits, inp, w['conc_type'], w['ncmp'], w['nppf'])
    assert np.array_equal(r['targets'].numpy(), tg)
    assert tg.sum() > 0, 'fixture must contain positive targets'
    assert abs(float(r['loss']) - loss) <= 1e-6 * abs(loss)


def test_oracle_loss_without_groundable_arguments_is_the_plain_mean():
    w, inp, logits, _, _ = _case('spat_gt5')
    inp['srl_arg_boxes_mask'] = torch.zeros_like(inp['srl_arg_boxes_mask'])        # code/mdl_conc_single.py:409-413
    r = vo.loss_forward(logits, inp, w['conc_type'], w['ncmp'], w['nppf'])
    tot = torch.nn.functional.binary_cross_entropy_with_logits(logits, r['targets'].float(), reduction='none')
    assert torch.allclose(r['loss'], tot.mean() * logits.shape[-1])


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_loss_matches_reference(name):
    w, inp, logits, loss, tg = _case(name)
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    dinp = {k: v.cuda() for k, v in inp.items()}
    with torch.no_grad():
        out = fn({'mdl_outs': logits.cuda()}, dinp)
        got_t = fn.compute_loss_targets(dinp)['targets_one']
    torch.cuda.synchronize()
    assert set(out) == {'loss', 'mdl_out_loss'} and out['loss'].shape == ()
    assert np.array_equal(got_t.cpu().numpy(), tg), 'IoU targets must be bit-exact'
    assert abs(float(out['loss']) - loss) <= 2e-6 * abs(loss), (float(out['loss']), loss)
    assert float(out['mdl_out_loss']) == float(out['loss'])


@pytest.mark.gpu
def test_cuda_loss_edge_cases():
    """no groundable argument -> plain mean; masked-out videos; zero-area proposals (-1 overlap rule)."""
    w, inp, logits, _, _ = _case('temp_gt5')
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    variants = []
    a = {k: v.clone() for k, v in inp.items()}
    a['srl_arg_boxes_mask'] = torch.zeros_like(a['srl_arg_boxes_mask'])
    variants.append(a)
    b = {k: v.clone() for k, v in inp.items()}
    b['num_cmp_msk'][:, 1:3] = 0
    variants.append(b)
    c = {k: v.clone() for k, v in inp.items()}
    c['pad_proposals'][:, ::7, 2] = c['pad_proposals'][:, ::7, 0]          # x2 == x1 and y2 == y1: zero-area anchors
    c['pad_proposals'][:, ::7, 3] = c['pad_proposals'][:, ::7, 1]
    c['pad_gt_bboxs'][:, 3] = 0                                             # a zero-area gt box that arguments point at
    variants.append(c)
    for v in variants:
        ref = vo.loss_forward(logits, v, w['conc_type'], w['ncmp'], w['nppf'])
        with torch.no_grad():
            dv = {k: t.cuda() for k, t in v.items()}
            out = fn({'mdl_outs': logits.cuda()}, dv)
            got_t = fn.compute_loss_targets(dv)['targets_one']
        assert torch.equal(got_t.cpu(), ref['targets'])
        assert abs(float(out['loss']) - float(ref['loss'])) <= 2e-6 * abs(float(ref['loss']))


def _golden_grad(name):
    gl = np.load(os.path.join(GOLD, f'loss_{name}.npz'))
    return gl['grad_sample'], float(gl['grad_sum']), float(gl['grad_abs_sum'])


def _check_grad(g, name):
    sample, gsum, gabs = _golden_grad(name)
    g = g.detach().cpu().reshape(-1)
    scale = float(np.abs(sample).max())
    assert np.abs(g[::7].numpy() - sample).max() <= 2e-6 * scale
    assert abs(float(g.double().sum()) - gsum) <= 1e-5 * gabs
    assert abs(float(g.double().abs().sum()) - gabs) <= 1e-5 * gabs


@pytest.mark.parametrize('name', NAMES)
def test_oracle_loss_gradient_matches_reference(name):
    """d loss / d logits (first link): autograd through the oracle restatement against the
    gradient autograd derives for the unmodified reference loss."""
    w, inp, logits, _, _ = _case(name)
    x = logits.clone().requires_grad_(True)
    vo.loss_forward(x, inp, w['conc_type'], w['ncmp'], w['nppf'])['loss'].backward()
    _check_grad(x.grad, name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_loss_gradient_matches_reference(name):
    w, inp, logits, loss, _ = _case(name)
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    x = logits.cuda().requires_grad_(True)
    out = fn({'mdl_outs': x}, {k: v.cuda() for k, v in inp.items()})
    assert abs(float(out['loss'].detach()) - loss) <= 2e-6 * abs(loss)
    (out['loss'] * 0.5).backward()                                     # a non-trivial upstream gradient
    assert x.grad.shape == x.shape
    _check_grad(x.grad * 2.0, name)


@pytest.mark.gpu
def test_cuda_loss_gradient_plain_mean_branch():
    """no groundable argument -> plain mean: every element gets weight P / n."""
    w, inp, logits, _, _ = _case('temp_gt5')
    inp = dict(inp)
    inp['srl_arg_boxes_mask'] = torch.zeros_like(inp['srl_arg_boxes_mask'])
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    x = logits.cuda().requires_grad_(True)
    fn({'mdl_outs': x}, {k: v.cuda() for k, v in inp.items()})['loss'].backward()
    xr = logits.clone().requires_grad_(True)
    vo.loss_forward(xr, inp, w['conc_type'], w['ncmp'], w['nppf'])['loss'].backward()
    assert torch.allclose(x.grad.cpu(), xr.grad, rtol=1e-5, atol=1e-9)


# ---------------------------------------------------------------------------------------------
# LossB_SEP (code/mdl_conc_sep.py:219-447)
# ---------------------------------------------------------------------------------------------
def _case_sep(name):
    w, batch = synth.workload(name)
    g = np.load(os.path.join(GOLD, f'{name}.npz'))
    gl = np.load(os.path.join(GOLD, f'loss_{name}.npz'))
    inp = dict(batch)
    inp.update(synth.make_loss_inputs_sep(batch, **w))
    tg = np.unpackbits(gl['targets'])[:int(np.prod(gl['targets_shape']))].reshape(gl['targets_shape']).astype(bool)
    out = {'mdl_outs': torch.from_numpy(g['mdl_outs']), 'vidf_outs': torch.from_numpy(g['vidf_outs'])}
    return w, inp, out, float(gl['loss']), float(gl['verb_loss']), tg


@pytest.mark.parametrize('name', list(synth.WORKLOADS_SEP))
def test_oracle_sep_loss_matches_reference(name):
    w, inp, out, loss, verb, tg = _case_sep(name)
    r = vo.loss_forward_sep(out, inp)
    assert np.array_equal(r['targets'].numpy(), tg) and tg.sum() > 0
    assert abs(float(r['loss']) - loss) <= 1e-6 * abs(loss)
    assert abs(float(r['verb_loss']) - verb) <= 1e-6 * abs(verb)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(synth.WORKLOADS_SEP))
def test_cuda_sep_loss_matches_reference(name):
    w, inp, out, loss, verb, tg = _case_sep(name)
    cfg, comm = synth.default_cfg('sep'), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    dinp = {k: v.cuda() for k, v in inp.items()}
    with torch.no_grad():
        res = fn({k: v.cuda() for k, v in out.items()}, dinp)
        got_t = fn.compute_loss_targets(dinp)['targets_one']
    torch.cuda.synchronize()
    assert set(res) == {'loss', 'mdl_out_loss', 'verb_loss'}
    assert np.array_equal(got_t.cpu().numpy(), tg), 'IoU targets must be bit-exact'
    assert abs(float(res['loss']) - loss) <= 2e-6 * abs(loss), (float(res['loss']), loss)
    assert abs(float(res['verb_loss']) - verb) <= 2e-6 * abs(verb), (float(res['verb_loss']), verb)


@pytest.mark.gpu
def test_cuda_sep_loss_edge_cases():
    """no groundable argument -> mean of the video-masked losses over ALL elements (code/mdl_conc_sep.py:355-363);
    single sentence slot expanded over the videos; no verb target at all -> NaN like the reference's empty mean."""
    w, inp, out, _, _, _ = _case_sep('sep_gt5')
    cfg, comm = synth.default_cfg('sep'), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    dout = {k: v.cuda() for k, v in out.items()}
    a = {k: v.clone() for k, v in inp.items()}
    a['srl_arg_boxes_mask'] = torch.zeros_like(a['srl_arg_boxes_mask'])
    b = {k: v.clone() for k, v in inp.items()}
    for k in ('srl_boxes', 'srl_boxes_lens', 'srl_arg_boxes_mask'):
        b[k] = b[k][:, :1].contiguous()
    for v in (a, b):
        ref = vo.loss_forward_sep(out, v)
        with torch.no_grad():
            got = fn(dout, {k: t.cuda() for k, t in v.items()})
            tg = fn.compute_loss_targets({k: t.cuda() for k, t in v.items()})['targets_one']
        assert torch.equal(tg.cpu(), ref['targets'])
        assert abs(float(got['loss']) - float(ref['loss'])) <= 2e-6 * abs(float(ref['loss']))
    c = {k: v.clone() for k, v in inp.items()}
    c['verb_cross_cmp_msk'] = torch.zeros_like(c['verb_cross_cmp_msk'])
    with torch.no_grad():
        got = fn(dout, {k: t.cuda() for k, t in c.items()})
    assert torch.isnan(got['verb_loss']) and torch.isnan(vo.loss_forward_sep(out, c)['verb_loss'])


@pytest.mark.parametrize('name', list(synth.WORKLOADS_SEP))
def test_oracle_sep_loss_gradient_matches_reference(name):
    w, inp, out, _, _, _ = _case_sep(name)
    o = dict(out)
    o['mdl_outs'] = out['mdl_outs'].clone().requires_grad_(True)
    vo.loss_forward_sep(o, inp)['loss'].backward()
    _check_grad(o['mdl_outs'].grad, name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(synth.WORKLOADS_SEP))
def test_cuda_sep_loss_gradient_matches_reference(name):
    w, inp, out, loss, _, _ = _case_sep(name)
    cfg, comm = synth.default_cfg('sep'), synth.default_comm(w['nppf'])
    fn = vb.get_mdl_loss_eval(cfg)['loss'](cfg, comm)
    x = out['mdl_outs'].cuda().requires_grad_(True)
    res = fn({'mdl_outs': x, 'vidf_outs': out['vidf_outs'].cuda()}, {k: v.cuda() for k, v in inp.items()})
    assert abs(float(res['loss'].detach()) - loss) <= 2e-6 * abs(loss)
    res['loss'].backward()
    _check_grad(x.grad, name)
