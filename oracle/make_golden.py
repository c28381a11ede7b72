"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) on the synthetic workloads of
vognet_pytorch_b200/synth.py.  Run in the build container:

    python oracle/make_golden.py            # all fixtures
    python oracle/make_golden.py spat_gt5   # one

Weights and inputs are NOT stored: they are regenerated bit-identically from (name, seed) by
synth.py; the fixtures hold only what the reference computed from them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                      # noqa: E402
from vognet_pytorch_b200 import synth                      # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def operator_case(name, d, n_heads, n_layers, Bt, N, rel, seed=3):
    """RelTransformer / Transformer called directly with a dense [Bt,N,N,H] bias tensor
    (operator-level boundary, code/transformer_code.py:244-279)."""
    tc = rh.reference_transformers()
    sd = synth.make_operator_state_dict(d, n_layers, seed=seed)
    x, pe = synth.make_operator_inputs(d, n_heads, Bt, N, seed=seed)
    if rel:
        m = tc.RelTransformer(d, 0, 0, d_hidden=d // 2, n_layers=n_layers, n_heads=n_heads,
                              drop_ratio=0.2, pe=False, d_pe=5)
    else:
        m = tc.Transformer(d, 0, 0, d_hidden=d // 2, n_layers=n_layers, n_heads=n_heads,
                           drop_ratio=0.2, pe=False)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        y = m(x, pe) if rel else m(x)
    np.savez(os.path.join(GOLD, f'op_{name}.npz'), y=y.numpy(),
             meta=np.array([d, n_heads, n_layers, Bt, N, int(rel), seed]))
    print(f'op_{name}: y {tuple(y.shape)} std {y.std():.3f}')


def model_case(name):
    w, batch = synth.workload(name)
    sd = synth.make_state_dict()
    mdl = rh.build_reference_model(w['conc_type'], w['nppf'], sd)
    keep = {}

    def hook(key):
        def f(mod, args, out):
            keep[key + '_in'] = args[0].detach().clone()
            keep[key + '_out'] = out.detach().clone()
        return f
    mdl.obj_txf.register_forward_hook(hook('obj'))
    mdl.mult_txf.register_forward_hook(hook('mul'))
    orig = mdl.retrieve_srl_arg_from_lang_encode

    def lang_hook(*a, **k):
        r = orig(*a, **k)
        keep['lang'] = r.detach().clone()
        return r
    mdl.retrieve_srl_arg_from_lang_encode = lang_hook

    with torch.no_grad():
        out = mdl(synth.clone_batch(batch))
        ev = rh.build_reference_evaluator(w['conc_type'], w['nppf'], w['ncmp'])
        sel = ev.get_out_results_boxes(out, batch)
    save = {
        'mdl_outs': out['mdl_outs'].numpy(),
        'mdl_outs_eval': out['mdl_outs_eval'].numpy(),
        'boxes': sel['boxes'].contiguous().numpy(),
        'scores': sel['scores'].contiguous().numpy(),
        'indexs': sel['indexs'].contiguous().numpy().astype(np.int64),
        'lang': keep['lang'].numpy(),
    }
    # intermediates: everything for the small case, a strided row sample otherwise
    stride = 1 if name == 'cpu_ref' else (23 if w['nppf'] == 5 else 397)
    for k in ('obj_out', 'mul_out'):
        t = keep[k]
        save[k + '_rows'] = t.reshape(-1, t.shape[-1])[::stride].contiguous().numpy()
        save[k + '_shape'] = np.array(t.shape)
    save['row_stride'] = np.array(stride)
    np.savez(os.path.join(GOLD, f'{name}.npz'), **save)
    sz = os.path.getsize(os.path.join(GOLD, f'{name}.npz')) / 1e6
    print(f'{name}: logits {tuple(out["mdl_outs"].shape)} std {out["mdl_outs"].std():.3f} '
          f'obj {tuple(keep["obj_out"].shape)} mul {tuple(keep["mul_out"].shape)}  {sz:.2f} MB')


def sep_case(name):
    """VOG_SEP + EvaluatorSEP of the unmodified reference (code/mdl_conc_sep.py:131-217,
    code/eval_vsrl_corr.py:162-220) -> tests/golden/{name}.npz."""
    w, batch = synth.workload(name)
    sd = synth.make_state_dict()
    mdl = rh.build_reference_model('sep', w['nppf'], sd)
    with torch.no_grad():
        out = mdl(synth.clone_batch(batch))
        ev = rh.build_reference_evaluator('sep', w['nppf'], w['ncmp'])
        sel = ev.get_out_results_boxes(out, batch)
    save = {k: out[k].contiguous().numpy() for k in ('mdl_outs', 'mdl_outs_eval', 'vidf_outs', 'fin_scores_loss', 'fin_scores')}
    save.update(boxes=sel['boxes'].contiguous().numpy(), scores=sel['scores'].contiguous().numpy(),
                indexs=sel['indexs'].contiguous().numpy().astype(np.int64))
    np.savez(os.path.join(GOLD, f'{name}.npz'), **save)
    print(f'{name}: logits {tuple(out["mdl_outs"].shape)} std {out["mdl_outs"].std():.3f} '
          f'fin_scores {out["fin_scores"].numpy().round(4).tolist()}')


GRAD_STRIDE = 7


def _reference_logit_grad(loss_fn, out, inp):
    """d loss / d mdl_outs through torch autograd on the UNMODIFIED reference loss: every GRAD_STRIDE-th element of
    the flattened gradient plus its float64 sum and absolute sum."""
    o = {k: v.clone() for k, v in out.items()}
    o['mdl_outs'] = o['mdl_outs'].clone().requires_grad_(True)
    res = loss_fn(o, {k: v.clone() for k, v in inp.items()})
    res['loss'].backward()
    g = o['mdl_outs'].grad.reshape(-1)
    return {'grad_sample': g[::GRAD_STRIDE].contiguous().numpy(), 'grad_sum': np.array(g.double().sum().item()),
            'grad_abs_sum': np.array(g.double().abs().sum().item())}


def param_grad_case(name='cpu_ref'):
    """d loss / d parameter for EVERY parameter of the unmodified reference (eval mode: dropout off, so the result is
    deterministic), through its own forward + LossB_SPAT and torch autograd -> tests/golden/grad_{name}.npz: per
    parameter the float64 L2 norm, sum and the first 8 elements.  Pins the oracle for the backward row."""
    w, batch = synth.workload(name)
    sd = synth.make_state_dict()
    mdl = rh.build_reference_model(w['conc_type'], w['nppf'], sd)
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    loss_fn = rh.build_reference_loss(w['conc_type'], w['nppf'])
    out = mdl(synth.clone_batch(inp))
    loss = loss_fn(out, {k: v.clone() for k, v in inp.items()})['loss']
    loss.backward()
    save = {'loss': loss.detach().numpy()}
    unused = []
    for k, p in mdl.named_parameters():
        if p.grad is None:
            unused.append(k)
            continue
        g = p.grad.double().reshape(-1)
        save['norm/' + k] = np.array(g.norm().item())
        save['sum/' + k] = np.array(g.sum().item())
        save['head/' + k] = g[:8].numpy()
    save['unused'] = np.array(unused)
    np.savez(os.path.join(GOLD, f'grad_{name}.npz'), **save)
    print(f'grad_{name}: loss {float(loss):.6f}, {len(save) // 3} parameters with gradients, unused: {unused}')


def ablation_case(name, mdl_name):
    """ImgGrnd_* / VidGrnd_* (mdl.name 'igrnd' / 'vgrnd', code/mdl_vog.py:400-410,526-535) of the unmodified reference
    on the workload `name` -> tests/golden/{mdl_name}_{name}.npz."""
    w, batch = synth.workload(name)
    mdl = rh.build_reference_model(w['conc_type'], w['nppf'], synth.make_state_dict(), mdl_name=mdl_name)
    with torch.no_grad():
        out = mdl(synth.clone_batch(batch))
    np.savez(os.path.join(GOLD, f'{mdl_name}_{name}.npz'), mdl_outs=out['mdl_outs'].numpy(),
             mdl_outs_eval=out['mdl_outs_eval'].numpy())
    print(f'{mdl_name}_{name}: logits {tuple(out["mdl_outs"].shape)} std {out["mdl_outs"].std():.3f}')


CFG_VARIANTS = {          # model-level configuration surface beyond the BASELINE settings (B=2 of spat/gt5, temp/gt5)
    'onefrm_spat': dict(conc='spat', cfg=dict(obj_one_frm=True), sd={}),
    'onefrm_temp': dict(conc='temp', cfg=dict(obj_one_frm=True), sd={}),
    'norel_spat': dict(conc='spat', cfg=dict(use_rel=False), sd={}),
    'l3h6_spat': dict(conc='spat', cfg=dict(n_layers=3, n_heads=6), sd=dict(n_layers_obj=3, n_layers_mul=3, n_heads=6)),
}


def cfg_variant_case(tag):
    """cfg.mdl.obj_tx.one_frm / use_rel=False / 3 layers x 6 heads (EXPTS.md:186-189) through the unmodified reference
    -> tests/golden/cfgvar_{tag}.npz."""
    v = CFG_VARIANTS[tag]
    batch = synth.make_batch(v['conc'], B=2, ncmp=4, nppf=5, seed=9)
    sd = synth.make_state_dict(seed=4, **v['sd'])
    mdl = rh.build_reference_model(v['conc'], 5, sd, **v['cfg'])
    with torch.no_grad():
        out = mdl(synth.clone_batch(batch))
    np.savez(os.path.join(GOLD, f'cfgvar_{tag}.npz'), mdl_outs=out['mdl_outs'].numpy(), mdl_outs_eval=out['mdl_outs_eval'].numpy())
    print(f'cfgvar_{tag}: logits std {out["mdl_outs"].std():.3f}')


ABLATIONS = [('spat_gt5', 'igrnd'), ('spat_gt5', 'vgrnd'), ('temp_gt5', 'igrnd'), ('temp_gt5', 'vgrnd')]


def loss_sep_case(name):
    """LossB_SEP of the unmodified reference on the golden SEP outputs -> tests/golden/loss_{name}.npz."""
    w, batch = synth.workload(name)
    gold = np.load(os.path.join(GOLD, f'{name}.npz'))
    out = {'mdl_outs': torch.from_numpy(gold['mdl_outs']), 'vidf_outs': torch.from_numpy(gold['vidf_outs'])}
    inp = dict(batch)
    inp.update(synth.make_loss_inputs_sep(batch, **w))
    loss_fn = rh.build_reference_loss('sep', w['nppf'])
    with torch.no_grad():
        res = loss_fn(out, {k: v.clone() for k, v in inp.items()})
        tg = loss_fn.compute_loss_targets({k: v.clone() for k, v in inp.items()})['targets_one']
    grad = _reference_logit_grad(loss_fn, out, inp)
    np.savez(os.path.join(GOLD, f'loss_{name}.npz'), loss=res['loss'].numpy(), mdl_out_loss=res['mdl_out_loss'].numpy(),
             verb_loss=res['verb_loss'].numpy(), **grad, targets=np.packbits(tg.numpy().astype(np.uint8)),
             targets_shape=np.array(tg.shape))
    print(f'loss_{name}: loss {float(res["loss"]):.6f} verb_loss {float(res["verb_loss"]):.6f} '
          f'positives {int(tg.sum())} of {tg.numel()}')


def relayout_case(name):
    """The reference's own host-side concatenation helpers (lifted from code/dat_loader_simple.py by
    ref_harness.reference_item_closures) on the per-video batch -> tests/golden/relayout_{name}.npz: the full
    proposal tensors and, for the wide feature tensors, every row's first 4 columns + float64 row sum (which pin
    the row permutation without storing megabytes)."""
    w, batch = synth.workload(name)
    save = {}
    for conc in ('spat', 'temp'):
        f, s_, p = rh.reference_concat_videos(batch, conc, synth.NFRM0, w['nppf'])
        save[f'{conc}_props'] = p.numpy()
        save[f'{conc}_feat_head'] = f[..., :4].contiguous().numpy()
        save[f'{conc}_feat_sum'] = f.double().sum(-1).numpy()
        save[f'{conc}_seg_head'] = s_[..., :4].contiguous().numpy()
        save[f'{conc}_seg_sum'] = s_.double().sum(-1).numpy()
    np.savez_compressed(os.path.join(GOLD, f'relayout_{name}.npz'), **save)
    print(f'relayout_{name}: spat props {save["spat_props"].shape} max x {save["spat_props"][..., 2].max():.1f} '
          f'temp max frame {save["temp_props"][..., 4].max():.0f}')


def loss_case(name):
    """LossB_SPAT / LossB_TEMP of the unmodified reference on the golden logits of `name` and the synthetic
    loss inputs -> tests/golden/loss_{name}.npz (loss value, boolean targets packed as uint8)."""
    w, batch = synth.workload(name)
    gold = np.load(os.path.join(GOLD, f'{name}.npz'))
    out = {'mdl_outs': torch.from_numpy(gold['mdl_outs'])}
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    loss_fn = rh.build_reference_loss(w['conc_type'], w['nppf'])
    with torch.no_grad():
        res = loss_fn(out, {k: v.clone() for k, v in inp.items()})
        tg = loss_fn.compute_loss_targets({k: v.clone() for k, v in inp.items()})['targets_one']
    grad = _reference_logit_grad(loss_fn, out, inp)
    np.savez(os.path.join(GOLD, f'loss_{name}.npz'), loss=res['loss'].numpy(), mdl_out_loss=res['mdl_out_loss'].numpy(),
             targets=np.packbits(tg.numpy().astype(np.uint8)), targets_shape=np.array(tg.shape), **grad)
    print(f'loss_{name}: loss {float(res["loss"]):.6f}  positives {int(tg.sum())} of {tg.numel()}')


# (weight seed, input seed, feature scale) combinations of the parity sweep: scores of the
# unmodified reference for every BASELINE configuration under different weights, inputs and feature magnitudes
SWEEP = {
    'spat_gt5': [(0, 1, 1.0), (1, 2, 1.0), (2, 3, 0.5), (3, 4, 2.0), (4, 5, 1.0), (5, 6, 0.5), (6, 7, 2.0), (7, 8, 1.0)],
    'temp_gt5': [(0, 1, 1.0), (1, 2, 1.0), (2, 3, 0.5), (3, 4, 2.0), (4, 5, 1.0), (5, 6, 0.5), (6, 7, 2.0), (7, 8, 1.0)],
    'spat_p100': [(1, 2, 1.0), (2, 3, 0.5), (3, 4, 2.0), (4, 5, 1.0)],
}


def sweep_batch(name, iseed, fscale):
    """the workload `name` with another input seed and the visual features scaled by `fscale`"""
    w, batch = synth.workload(name, seed=iseed)
    batch = dict(batch)
    for k in ('pad_region_feature', 'seg_feature_for_frms'):
        batch[k] = batch[k] * fscale
    return w, batch


def sweep_case(name):
    """-> tests/golden/sweep_{name}.npz: masked scores of the unmodified reference for every combination of SWEEP
    (fp16 storage would lose the 1e-3 tolerance: kept as fp32, compressed)."""
    save = {'combos': np.array(SWEEP[name], dtype=np.float64)}
    for i, (ws, iseed, fs) in enumerate(SWEEP[name]):
        w, batch = sweep_batch(name, iseed, fs)
        mdl = rh.build_reference_model(w['conc_type'], w['nppf'], synth.make_state_dict(seed=ws))
        with torch.no_grad():
            out = mdl(synth.clone_batch(batch))
        save[f'scores_{i}'] = out['mdl_outs_eval'].numpy()
        print(f'sweep_{name}[{i}] weights {ws} inputs {iseed} x{fs}: logits std {out["mdl_outs"].std():.3f} '
              f'score range {float(out["mdl_outs_eval"].min()):.3f}..{float(out["mdl_outs_eval"].max()):.3f}')
    np.savez_compressed(os.path.join(GOLD, f'sweep_{name}.npz'), **save)


def operator_case_n2000(name='rel_d768_h3_l1_n2000', d=768, n_heads=3, Bt=2, N=2000, seed=5, stride=16):
    """The operator-level boundary at the north-star sequence length: RelTransformer(x, x_pe) with the reference's
    DENSE x_pe [Bt,N,N,H] (96 MB) -> every `stride`-th output row."""
    tc = rh.reference_transformers()
    sd = synth.make_operator_state_dict(d, 1, seed=seed)
    x, pe = synth.make_operator_inputs(d, n_heads, Bt, N, seed=seed)
    m = tc.RelTransformer(d, 0, 0, d_hidden=d // 2, n_layers=1, n_heads=n_heads, drop_ratio=0.2, pe=False, d_pe=5)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with torch.no_grad():
        y = m(x, pe)
    np.savez_compressed(os.path.join(GOLD, f'op_{name}.npz'), y_rows=y.reshape(-1, d)[::stride].contiguous().numpy(),
                        meta=np.array([d, n_heads, 1, Bt, N, 1, seed, stride]))
    print(f'op_{name}: y {tuple(y.shape)} std {y.std():.3f}, {y.reshape(-1, d)[::stride].shape[0]} rows kept')


OPS = {
    'rel_d512_h3_l2':  dict(d=512, n_heads=3, n_layers=2, Bt=2, N=37, rel=True),
    'rel_d768_h3_l1':  dict(d=768, n_heads=3, n_layers=1, Bt=3, N=50, rel=True),
    'rel_d512_h6_l1':  dict(d=512, n_heads=6, n_layers=1, Bt=1, N=130, rel=True),   # EXPTS.md:186-189 ablation
    'plain_d512_h3_l1': dict(d=512, n_heads=3, n_layers=1, Bt=2, N=64, rel=False),
}

if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    want = sys.argv[1:]
    for nm, kw in OPS.items():
        if not want or ('op_' + nm) in want:
            operator_case(nm, **kw)
    for nm in synth.WORKLOADS:
        if not want or nm in want:
            model_case(nm)
    for nm in synth.WORKLOADS:
        if not want or ('loss_' + nm) in want:
            loss_case(nm)
    for nm in synth.WORKLOADS_SEP:
        if not want or nm in want:
            sep_case(nm)
    for nm in synth.WORKLOADS_SEP:
        if not want or ('loss_' + nm) in want:
            loss_sep_case(nm)
    for nm in synth.WORKLOADS_SEP:
        if not want or ('relayout_' + nm) in want:
            relayout_case(nm)
    for nm, mn in ABLATIONS:
        if not want or f'{mn}_{nm}' in want:
            ablation_case(nm, mn)
    for tag in CFG_VARIANTS:
        if not want or f'cfgvar_{tag}' in want:
            cfg_variant_case(tag)
    if not want or 'grad_cpu_ref' in want:
        param_grad_case('cpu_ref')
    for nm in SWEEP:
        if not want or ('sweep_' + nm) in want:
            sweep_case(nm)
    if not want or 'op_rel_d768_h3_l1_n2000' in want:
        operator_case_n2000()
