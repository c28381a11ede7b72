"""Parity margins over a sweep of (weight seed, input seed, feature scale) per BASELINE configuration (VERDICT r1
item 6): the masked scores of the CUDA path against the UNMODIFIED reference's (tests/golden/sweep_*.npz, written by
oracle/make_golden.py), in the compute mode each configuration is specified for, plus the exact-fp32 mode.
Reports max |d score| and the selection flips (all / inside wide top-2 gaps) per combination and writes them to
gpurun_out/parity_margins.json (copied to profiles/r2/ after a GPU run)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import vognet_pytorch_b200 as vb              # noqa: E402
from vognet_pytorch_b200 import ops, synth     # noqa: E402

DEV = 'cuda:0'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {'fp32x': 1e-4, 'tf32': 1e-3, 'bf16': 1e-2}
MODES = {'spat_gt5': ('tf32', 'fp32x', 'bf16'), 'temp_gt5': ('tf32', 'fp32x', 'bf16'), 'spat_p100': ('bf16', 'tf32')}
_RESULTS = {}


def _flips(w, ref_scores, got_scores, tol):
    """argmax-over-proposals flips per (query, srl, video, frame) group and how many sit in groups whose reference
    top-2 gap exceeds 2 x tol (those must not flip)."""
    B, _, nsrl, P = ref_scores.shape
    nppf = w['nppf']
    r = torch.from_numpy(ref_scores).view(B, nsrl, -1, nppf)
    g = torch.from_numpy(got_scores).view(B, nsrl, -1, nppf)
    top2 = r.topk(2, -1).values
    wide = (top2[..., 0] - top2[..., 1]) > 2 * tol
    flip = r.argmax(-1) != g.argmax(-1)
    return int(flip.sum()), int((flip & wide).sum()), int(wide.sum()), flip.numel()


@pytest.mark.parametrize('name', ['spat_gt5', 'temp_gt5', 'spat_p100'])
def test_parity_sweep(golden, name):
    g = golden('sweep_' + name)
    rows = []
    for i, (ws, iseed, fs) in enumerate(g['combos']):
        w, batch = synth.workload(name, seed=int(iseed))
        batch = dict(batch)
        for k in ('pad_region_feature', 'seg_feature_for_frms'):
            batch[k] = batch[k] * float(fs)
        cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
        mdl = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
        mdl.load_state_dict(synth.make_state_dict(seed=int(ws)), strict=True)
        mdl = mdl.to(DEV).eval()
        dbatch = synth.clone_batch(batch, DEV)
        ref = g[f'scores_{i}']
        for mode in MODES[name]:
            mdl.set_compute(mode)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            mdl(dbatch)
            t0.record()
            got = mdl(dbatch)['mdl_outs_eval']
            t1.record()
            torch.cuda.synchronize()
            got = got.cpu().numpy()
            err = float(np.abs(got - ref).max())
            nflip, nflip_wide, nwide, ngroups = _flips(w, ref, got, TOL[mode])
            rows.append(dict(weights_seed=int(ws), input_seed=int(iseed), feature_scale=float(fs), compute=mode,
                             max_abs_dscore=err, tol=TOL[mode], margin=TOL[mode] / max(err, 1e-30),
                             selection_flips=nflip, flips_in_wide_gap_groups=nflip_wide, wide_gap_groups=nwide,
                             groups=ngroups, eager_ms=t0.elapsed_time(t1)))
            assert np.isfinite(got).all()
            assert err < TOL[mode], (name, mode, ws, iseed, fs, err)
            assert nflip_wide == 0, (name, mode, nflip_wide)
        del mdl
    _RESULTS[name] = rows
    worst = {}
    for r in rows:
        worst[r['compute']] = max(worst.get(r['compute'], 0.0), r['max_abs_dscore'])
    print(f'\n[sweep {name}] worst max|dscore| per mode: ' + ', '.join(f'{k} {v:.2e} (tol {TOL[k]:g})' for k, v in worst.items()))
    out_dir = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out_dir, exist_ok=True)
        path = os.path.join(out_dir, 'parity_margins.json')
        prev = json.load(open(path)) if os.path.exists(path) else {}
        prev[name] = rows
        json.dump(prev, open(path, 'w'), indent=1)
    except OSError:
        pass


def test_operator_dense_bias_at_n2000(golden):
    """RelTransformer(x, x_pe) with the reference's DENSE [Bt,N,N,H] bias tensor at N = 2000 (the multimodal sequence
    length of spat/p100): every 16th output row of the unmodified reference."""
    g = golden('op_rel_d768_h3_l1_n2000')
    d, H, L, Bt, N, rel, seed, stride = [int(v) for v in g['meta']]
    m = vb.RelTransformer(d, 0, 0, d_hidden=d // 2, n_layers=L, n_heads=H, drop_ratio=0.2, pe=False, d_pe=5)
    m.load_state_dict(synth.make_operator_state_dict(d, L, seed=seed), strict=True)
    x, pe = synth.make_operator_inputs(d, H, Bt, N, seed=seed)
    x, pe = x.to(DEV), pe.to(DEV)
    for mode, tol in (('fp32x', 2e-4), ('tf32', 1e-2), ('bf16', 4e-2)):
        mm = m.to(DEV).eval().set_compute(mode)
        with torch.no_grad():
            y = mm(x, pe)
        torch.cuda.synchronize()
        err = float(np.abs(y.reshape(-1, d)[::stride].cpu().numpy() - g['y_rows']).max())
        print(f'\n[op dense x_pe N={N} / {mode}] max|dy| {err:.2e}')
        assert err < tol, (mode, err)
