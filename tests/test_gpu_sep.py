"""SEP concatenation (code/mdl_conc_sep.py:13-217) on the GPU: VOG_SEP forward in every
compute mode and EvaluatorSEP against the golden vectors of the unmodified reference.

Tolerances: fp32x 1e-4 (exact fp32 arithmetic, different summation order), tf32 1e-3, bf16 1e-2 - the ones
BASELINE.json's north_star states for pred_scores."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import vognet_pytorch_b200 as vb              # noqa: E402
from vognet_pytorch_b200 import ops, synth     # noqa: E402
from oracle import vog_oracle as vo            # noqa: E402

DEV = 'cuda:0'
TOL = {'fp32x': 1e-4, 'tf32': 1e-3, 'bf16': 1e-2}
KEYS = ('mdl_outs_eval', 'vidf_outs', 'fin_scores_loss', 'fin_scores')


def _model(name):
    w, batch = synth.workload(name)
    cfg = synth.default_cfg('sep')
    comm = synth.default_comm(w['nppf'])
    sel = vb.get_mdl_loss_eval(cfg)
    mdl = sel['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    return w, batch, mdl.to(DEV).eval(), sel['eval'](cfg, comm, DEV)


@pytest.mark.parametrize('name', ['sep_gt5', 'sep_p100'])
@pytest.mark.parametrize('mode', ['fp32x', 'tf32', 'bf16'])
def test_sep_forward_golden(golden, name, mode):
    g = golden(name)
    w, batch, mdl, ev = _model(name)
    mdl.set_compute(mode)
    dbatch = synth.clone_batch(batch, DEV)
    out = mdl(dbatch)
    torch.cuda.synchronize()
    for k in KEYS:
        got = out[k].cpu().numpy()
        assert got.shape == g[k].shape, k
        assert np.isfinite(got).all(), k
        err = np.abs(got - g[k]).max()
        print(f'[{name}/{mode}] {k}: max|d| {err:.2e}')
        # vidf_outs is a logit (not squashed): same relative tolerance on a O(1) quantity
        assert err < TOL[mode], (k, err)
    # masked-out video: every output of it is exactly zero
    dead = (batch['num_cmp_msk'] == 0)
    assert (out['mdl_outs_eval'].cpu()[dead] == 0).all() and (out['fin_scores'].cpu()[dead] == 0).all()

    # evaluator: bit-exact on the product's own scores (checked against the oracle's selection of the same scores),
    # within tolerance of the golden scores
    sel = ev.get_out_results_boxes(out, dbatch)
    ref = vo.select_boxes_sep({k: out[k].cpu() for k in ('mdl_outs_eval', 'fin_scores')}, batch['pad_proposals'], w['nppf'])
    for k in ('boxes', 'scores', 'indexs'):
        assert torch.equal(sel[k].cpu(), ref[k]), k
    assert np.abs(sel['scores'].cpu().numpy() - g['scores']).max() < TOL[mode]
    gap = np.sort(g['fin_scores'], -1)
    if (gap[:, -1] - gap[:, -2] > 2 * TOL[mode]).all():
        assert np.array_equal(sel['indexs'].cpu().numpy(), g['indexs'])


def test_sep_graph_replay_matches_eager(golden):
    w, batch, mdl, _ = _model('sep_gt5')
    mdl.set_compute('tf32')
    dbatch = synth.clone_batch(batch, DEV)
    eager = {k: v.clone() for k, v in mdl(dbatch).items()}
    mdl.use_cuda_graph = True
    for _ in range(2):
        out = mdl(dbatch)
    torch.cuda.synchronize()
    for k in eager:
        assert torch.equal(out[k], eager[k]), k
    # a second batch through the same captured graph
    batch2 = synth.make_batch_sep(B=w['B'], ncmp=w['ncmp'], nppf=w['nppf'], seed=7)
    d2 = synth.clone_batch(batch2, DEV)
    out2 = {k: v.clone() for k, v in mdl(d2).items()}
    mdl.use_cuda_graph = False
    ref2 = mdl(d2)
    for k in ref2:
        assert torch.equal(out2[k], ref2[k]), k


def test_sep_single_sentence_slot_expands(golden):
    """One sentence slot for ncmp videos (code/mdl_conc_sep.py:165-173,190-195 expand 1 -> ncmp): same result as
    the batch with the sentence repeated."""
    w, batch, mdl, _ = _model('sep_gt5')
    mdl.set_compute('tf32')
    rep = dict(batch)
    for k in mdl._SEP_LANG_KEYS:
        rep[k] = batch[k][:, :1].expand_as(batch[k]).contiguous()
    one = dict(batch)
    for k in mdl._SEP_LANG_KEYS:
        one[k] = batch[k][:, :1].contiguous()
    a = mdl(synth.clone_batch(rep, DEV))
    b = mdl(synth.clone_batch(one, DEV))
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_sep_fin_scores_kernel_matches_torch():
    g = torch.Generator().manual_seed(3)
    Bq, nsrl, P1 = 7, 5, 130
    logits = torch.randn(Bq, nsrl, P1, generator=g)
    vidf = torch.randn(Bq, generator=g)
    msk = (torch.rand(Bq, nsrl, generator=g) > 0.3).long()
    msk[:, 0] = 1
    verb = torch.randint(0, nsrl, (Bq,), generator=g)
    cm = (torch.rand(Bq, generator=g) > 0.3).long()
    fl, fe = ops.sep_fin_scores(logits.to(DEV), vidf.to(DEV), msk.to(DEV), verb.to(DEV), cm.to(DEV))
    best = torch.sigmoid(logits).max(-1)[0].scatter(1, verb.view(Bq, 1), torch.sigmoid(vidf).view(Bq, 1)) * msk.float()
    assert torch.allclose(fl.cpu(), best * cm.view(Bq, 1).float(), atol=1e-6)
    assert torch.allclose(fe.cpu(), best.sum(-1) / msk.sum(-1).float() * cm.float(), atol=1e-6)


def test_sep_rejects_single_video_layout():
    _, batch, mdl, _ = _model('sep_gt5')
    _, flat = synth.workload('temp_gt5')
    with pytest.raises(ValueError):
        mdl(synth.clone_batch(flat, DEV))
