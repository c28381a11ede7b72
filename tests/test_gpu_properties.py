"""Size-independent properties of the forward fusion path, checked at the FULL BASELINE sizes (spat/p100: N = 4000 /
2000 attention) where the CPU oracle is too slow to be the checker:

  * queries are independent: a query scored inside a batch of 4 equals the same query scored alone
  * proposals of a (frame, video) slot are a SET: permuting them (features, boxes) permutes the scores the same way -
    nothing but the boxes carries position (code/mdl_vog.py:456-490), the regroup / un-regroup must round-trip
  * masks act on the outputs only: a masked video / argument slot is exactly zero, everything else is bit-identical
  * the evaluator's scores are the group maxima of mdl_outs_eval and its boxes are rows of pad_proposals (bit-exact)

Tolerances on the first two: a different batch size or proposal order changes the fp32 summation order inside split-K
GEMMs and attention tiles by ~1e-7, and such a last-bit difference can flip the bf16 rounding of an attention operand
(one Q/K/V element moves by 2^-8 relative) - measured 2.8e-4 on scores at spat/gt5.  The bound is therefore the same
order as the noise floor of the mode (half its tolerance), while any cross-talk or indexing bug moves scores by O(0.1)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import vognet_pytorch_b200 as vb              # noqa: E402
from vognet_pytorch_b200 import synth          # noqa: E402

DEV = 'cuda:0'
CASES = [('spat_gt5', 'tf32', 5e-4), ('spat_p100', 'bf16', 5e-3), ('temp_p100', 'bf16', 5e-3)]


def _model(name, mode):
    w, batch = synth.workload(name)
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    sel = vb.get_mdl_loss_eval(cfg)
    mdl = sel['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    return w, synth.clone_batch(batch, DEV), mdl.to(DEV).eval().set_compute(mode), sel['eval'](cfg, comm, DEV)


@pytest.mark.parametrize('name,mode,tol', CASES)
def test_queries_are_independent(name, mode, tol):
    w, batch, mdl, _ = _model(name, mode)
    full = mdl(batch)['mdl_outs_eval']
    one = mdl({k: v[2:3].contiguous() for k, v in batch.items()})['mdl_outs_eval']
    err = (full[2:3] - one).abs().max().item()
    print(f'\n[{name}/{mode}] query alone vs in a batch of {w["B"]}: max|d| {err:.2e}')
    assert err < tol


@pytest.mark.parametrize('name,mode,tol', CASES)
def test_proposals_of_a_slot_are_a_set(name, mode, tol):
    w, batch, mdl, _ = _model(name, mode)
    nppf = w['nppf']
    P = batch['pad_proposals'].shape[1]
    g = torch.Generator().manual_seed(5)
    perm = torch.cat([torch.randperm(nppf, generator=g) + s * nppf for s in range(P // nppf)]).to(DEV)
    pb = dict(batch)
    pb['pad_region_feature'] = batch['pad_region_feature'][:, perm].contiguous()
    pb['pad_proposals'] = batch['pad_proposals'][:, perm].contiguous()
    a = mdl(batch)['mdl_outs_eval']
    b = mdl(pb)['mdl_outs_eval']
    err = (a[..., perm] - b).abs().max().item()
    print(f'\n[{name}/{mode}] permuted proposals: max|d| {err:.2e}')
    assert err < tol
    assert (a - b).abs().max().item() > 10 * tol, 'the permutation must actually move scores'


@pytest.mark.parametrize('name,mode,tol', CASES)
def test_masks_touch_only_their_outputs(name, mode, tol):
    w, batch, mdl, _ = _model(name, mode)
    B, ncmp, nppf = w['B'], w['ncmp'], w['nppf']
    base = mdl(batch)
    mb = {k: v.clone() for k, v in batch.items()}
    mb['num_cmp_msk'][1, 2] = 0
    got = mdl(mb)
    assert torch.equal(got['mdl_outs'], base['mdl_outs'])                  # logits are not masked
    P = base['mdl_outs'].shape[-1]
    idx = torch.arange(P, device=DEV)
    vid = (idx // nppf) % ncmp if w['conc_type'] == 'spat' else idx // (P // ncmp)
    dead = torch.zeros(B, 1, 1, P, dtype=torch.bool, device=DEV)
    dead[1, 0, 0] = vid == 2
    dead = dead.expand_as(base['mdl_outs_eval'])
    assert (got['mdl_outs_eval'][dead] == 0).all()
    assert torch.equal(got['mdl_outs_eval'][~dead], base['mdl_outs_eval'][~dead])
    # an empty argument slot: exactly zero scores
    nsrl_valid = batch['srl_arg_inds_msk'].sum(-1)
    for b in range(B):
        n = int(nsrl_valid[b, 0])
        assert (base['mdl_outs_eval'][b, 0, n:] == 0).all()


@pytest.mark.parametrize('name,mode,tol', CASES)
def test_selection_is_the_group_maximum(name, mode, tol):
    w, batch, mdl, ev = _model(name, mode)
    B, ncmp, nppf = w['B'], w['ncmp'], w['nppf']
    out = mdl(batch)
    sel = ev.get_out_results_boxes(out, batch)
    s = out['mdl_outs_eval'][:, 0]
    nsrl = s.shape[1]
    if w['conc_type'] == 'spat':
        grp = s.view(B, nsrl, 10, ncmp, nppf).transpose(2, 3)
        props = batch['pad_proposals'].view(B, 10, ncmp, nppf, 7).transpose(1, 2)
    else:
        grp = s.view(B, nsrl, ncmp, 10, nppf)
        props = batch['pad_proposals'].view(B, ncmp, 10, nppf, 7)
    mx, am = grp.max(-1)
    assert torch.equal(sel['scores'], mx.contiguous())
    want = torch.gather(props.unsqueeze(1).expand(B, nsrl, ncmp, 10, nppf, 7), -2,
                        am[..., None, None].expand(B, nsrl, ncmp, 10, 1, 7)).squeeze(-2)
    assert torch.equal(sel['boxes'], want)
    if w['conc_type'] == 'spat':
        assert torch.equal(sel['indexs'], mx.argmax(2))
