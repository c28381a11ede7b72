"""Live check of the oracle against the UNMODIFIED reference (only where /root/reference exists,
i.e. in the build container; skipped on the GPU box).  Uses inputs and weights that are NOT in the
golden fixtures (other seeds, 3-layer / 6-head ablation config of EXPTS.md:186-189)."""
import pytest
import torch

from oracle import ref_harness as rh
from oracle import vog_oracle as vo
from vognet_pytorch_b200 import synth

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason='reference tree not present')


@pytest.mark.parametrize('conc,seed', [('spat', 11), ('temp', 12)])
def test_oracle_matches_live_reference(conc, seed):
    sd = synth.make_state_dict(seed=seed)
    batch = synth.make_batch(conc, B=2, ncmp=4, nppf=5, seed=seed)
    mdl = rh.build_reference_model(conc, 5, sd)
    with torch.no_grad():
        ref = mdl(synth.clone_batch(batch))
        got = vo.vog_forward(sd, batch, conc, 5)
        ev = rh.build_reference_evaluator(conc, 5, 4)
        sel_ref = ev.get_out_results_boxes(ref, batch)
        sel = vo.select_boxes(ref['mdl_outs_eval'], batch['pad_proposals'], conc, 4, 5)
    assert (got['mdl_outs'] - ref['mdl_outs']).abs().max() < 5e-5
    assert (got['mdl_outs_eval'] - ref['mdl_outs_eval']).abs().max() < 2e-5
    assert torch.equal(sel['boxes'], sel_ref['boxes'].contiguous())
    assert torch.equal(sel['scores'], sel_ref['scores'].contiguous())
    assert torch.equal(sel['indexs'].float(), sel_ref['indexs'].float().expand_as(sel['indexs']))


def test_oracle_matches_live_reference_3layer_6head():
    cfgkw = dict(n_layers=3, n_heads=6)
    sd = synth.make_state_dict(seed=5, n_layers_obj=3, n_layers_mul=3, n_heads=6)
    batch = synth.make_batch('spat', B=1, ncmp=4, nppf=5, seed=5)
    mdl = rh.build_reference_model('spat', 5, sd, **cfgkw)
    with torch.no_grad():
        ref = mdl(synth.clone_batch(batch))
        got = vo.vog_forward(sd, batch, 'spat', 5, n_heads=6)
    assert (got['mdl_outs'] - ref['mdl_outs']).abs().max() < 1e-4


def test_sep_oracle_matches_live_reference():
    sd = synth.make_state_dict(seed=21)
    batch = synth.make_batch_sep(B=2, ncmp=3, nppf=5, seed=21)
    mdl = rh.build_reference_model('sep', 5, sd)
    with torch.no_grad():
        ref = mdl(synth.clone_batch(batch))
        got = vo.vog_forward_sep(sd, synth.clone_batch(batch), 5)
        sel_ref = rh.build_reference_evaluator('sep', 5, 3).get_out_results_boxes(ref, batch)
        sel = vo.select_boxes_sep(ref, batch['pad_proposals'], 5)
    for k in ('mdl_outs', 'mdl_outs_eval', 'vidf_outs', 'fin_scores_loss', 'fin_scores'):
        assert (got[k] - ref[k]).abs().max() < 5e-5, k
    for k in ('boxes', 'scores', 'indexs'):
        assert torch.equal(sel[k], sel_ref[k].contiguous()), k


@pytest.mark.parametrize('conc', ['spat', 'temp'])
def test_concat_oracle_matches_live_reference_helpers(conc):
    batch = synth.make_batch_sep(B=3, ncmp=3, nppf=5, seed=31)
    ref = rh.reference_concat_videos(batch, conc, 10, 5)
    got = vo.concat_videos(batch['pad_region_feature'], batch['seg_feature_for_frms'], batch['pad_proposals'], conc, 10, 5)
    for r, g in zip(ref, got):
        assert torch.equal(r, g)
