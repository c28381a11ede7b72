"""CPU-side checks: the C-ABI library loads and exports every symbol include/vog_b200.h declares,
the module surface keeps the reference's checkpoint contract, and the selector mirrors
code/mdl_selector.py.  No GPU compute."""
import os
import re

import pytest
import torch

import vognet_pytorch_b200 as vb
from vognet_pytorch_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'vog_b200.h')).read()
    declared = set(re.findall(r'\b(vog_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), f'{name} declared in include/vog_b200.h but not exported'
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert L.vog_abi_version() >= 1
    assert L.vog_last_error() is not None


@pytest.mark.parametrize('conc', ['spat', 'temp', 'sep'])
def test_checkpoint_contract(conc):
    cfg, comm = synth.default_cfg(conc), synth.default_comm(5)
    sel = vb.get_mdl_loss_eval(cfg)
    mdl = sel['mdl'](cfg, comm)
    sd = synth.make_state_dict()
    assert list(mdl.state_dict().keys()) == list(sd.keys()) or set(mdl.state_dict()) == set(sd)
    mdl.load_state_dict(sd, strict=True)
    for k, v in mdl.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    assert sum(p.numel() for p in mdl.parameters()) == 45250727       # SURVEY.md section 2b [probed]
    assert not list(mdl.buffers())


def test_selector_surface():
    for conc, name, cls in (('spat', 'vog', 'VOG_SPAT'), ('temp', 'vog', 'VOG_TEMP'),
                            ('spat', 'vgrnd', 'VidGrnd_SPAT'), ('temp', 'igrnd', 'ImgGrnd_TEMP')):
        cfg = synth.default_cfg(conc)
        cfg.mdl.name = name
        out = vb.get_mdl_loss_eval(cfg)
        assert set(out) == {'mdl', 'loss', 'eval'}
        assert out['mdl'].__name__ == cls
    for conc in ('sep', 'svsq'):                     # code/mdl_selector.py:29 - both select the SEP classes
        cfg = synth.default_cfg(conc)
        out = vb.get_mdl_loss_eval(cfg)
        assert out['mdl'].__name__ == 'VOG_SEP' and out['eval'].__name__ == 'EvaluatorSEP'
        assert out['loss'].__name__ == 'LossB_SEP'
        assert out['loss'](cfg, synth.default_comm(5)).loss_keys == ['loss', 'mdl_out_loss', 'verb_loss']
    cfg = synth.default_cfg('spat')
    cfg.mdl.name = 'nope'
    with pytest.raises(NotImplementedError):
        vb.get_mdl_loss_eval(cfg)


def test_ablation_variants_have_reference_parameter_sets():
    cfg, comm = synth.default_cfg('spat'), synth.default_comm(5)
    from vognet_pytorch_b200 import mdl_vog
    ig = set(mdl_vog.ImgGrnd_SPAT(cfg, comm).state_dict())
    vg = set(mdl_vog.VidGrnd_SPAT(cfg, comm).state_dict())
    vo_ = set(mdl_vog.VOG_SPAT(cfg, comm).state_dict())
    assert not any(k.startswith(('obj_txf', 'mult_txf', 'pe_')) for k in ig)
    assert any(k.startswith('obj_txf') for k in vg) and not any(k.startswith('mult_txf') for k in vg)
    assert ig < vg < vo_


def test_forward_refuses_cpu():
    cfg, comm = synth.default_cfg('spat'), synth.default_comm(5)
    mdl = vb.VOG_SPAT(cfg, comm).eval()
    _, batch = synth.workload('cpu_ref')
    with pytest.raises(RuntimeError):
        mdl(batch)


def test_synth_is_deterministic():
    a, b = synth.make_state_dict(), synth.make_state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    _, b1 = synth.workload('spat_gt5')
    _, b2 = synth.workload('spat_gt5')
    assert all(torch.equal(b1[k], b2[k]) for k in b1)


def test_sep_batch_layout():
    """SEP batches carry an [ncmp] axis everywhere (code/mdl_conc_sep.py:131-160) and are the unfolded single-video
    pairs: video c of query b is pseudo-query b*ncmp+c."""
    w, batch = synth.workload('sep_gt5')
    B, ncmp, nppf = w['B'], w['ncmp'], w['nppf']
    assert tuple(batch['pad_region_feature'].shape) == (B, ncmp, 10 * nppf, 2048)
    assert tuple(batch['seg_feature_for_frms'].shape) == (B, ncmp, 10, 3072)
    assert tuple(batch['pad_proposals'].shape) == (B, ncmp, 10 * nppf, 7)
    assert tuple(batch['srl_arg_words_ind'].shape) == (B, ncmp, 5, 20)
    assert tuple(batch['verb_ind_in_srl'].shape) == (B, ncmp)
    assert (batch['verb_ind_in_srl'] < batch['srl_arg_inds_msk'].sum(-1)).all()
    assert batch['pad_proposals'][..., 4].max() == 9 and batch['pad_proposals'][..., 0].max() < 720
    assert batch['num_cmp_msk'][1, -1] == 0 and batch['num_cmp_msk'].sum() == B * ncmp - B // 2


def test_sep_flatten_is_the_single_video_batch():
    """VOG_SEP re-views the [B,ncmp,...] batch as B*ncmp single-video queries (host logic only, no kernels): the result
    equals the temp batch the SEP fixture was unfolded from, without copying the visual tensors; one sentence slot is
    expanded over the videos (code/mdl_conc_sep.py:165-173)."""
    w, batch = synth.workload('sep_gt5')
    B, ncmp, nppf = w['B'], w['ncmp'], w['nppf']
    cfg, comm = synth.default_cfg('sep'), synth.default_comm(nppf)
    mdl = vb.get_mdl_loss_eval(cfg)['mdl'](cfg, comm)
    flat, (b_, c_) = mdl._sep_flatten(batch)
    assert (b_, c_) == (B, ncmp)
    one = synth.make_batch(conc_type='temp', B=B * ncmp, ncmp=1, nppf=nppf, seed=1)
    for k in ('pad_region_feature', 'seg_feature_for_frms', 'pad_proposals', 'srl_arg_words_ind', 'srl_arg_word_mask',
              'srl_arg_word_mask_len', 'srl_arg_words_capture', 'srl_arg_inds_msk'):
        assert torch.equal(flat[k], one[k]), k
    assert flat['pad_region_feature'].data_ptr() == batch['pad_region_feature'].data_ptr()      # a view, not a copy
    assert tuple(flat['num_cmp_msk'].shape) == (B * ncmp, 1) and tuple(flat['verb_ind_in_srl'].shape) == (B * ncmp,)
    single = dict(batch)
    for k in mdl._SEP_LANG_KEYS:
        single[k] = batch[k][:, :1].contiguous()
    flat1, _ = mdl._sep_flatten(single)
    assert torch.equal(flat1['srl_arg_words_ind'].view(B, ncmp, 5, 20)[:, 3], batch['srl_arg_words_ind'][:, 0])
    bad = dict(batch)
    bad['pad_region_feature'] = batch['pad_region_feature'].reshape(B, -1, 2048)
    with pytest.raises(ValueError):
        mdl._sep_flatten(bad)


def test_ctypes_signatures_match_the_header():
    """ABI drift guard: for every declaration in include/vog_b200.h the ctypes binding has the same number of
    arguments, pointer arguments are bound as void*, and int / int64_t / float / double map to the matching ctypes."""
    import ctypes
    hdr = open(os.path.join(ROOT, 'include', 'vog_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', ' ', hdr, flags=re.S)
    decls = re.findall(r'\b(vog_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', hdr, flags=re.S)
    assert len(decls) >= 40
    ctype_of = {'int': ctypes.c_int, 'int64_t': ctypes.c_int64, 'uint64_t': ctypes.c_uint64, 'float': ctypes.c_float,
                'double': ctypes.c_double}
    for name, args in decls:
        args = ' '.join(args.split())
        params = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
        bound = _lib._SIGNATURES[name]
        assert len(params) == len(bound), (name, params, bound)
        for prm, ct in zip(params, bound):
            if '*' in prm:
                assert ct is ctypes.c_void_p, (name, prm, ct)
            else:
                base = prm.replace('const ', '').split()[0]
                assert ct is ctype_of[base], (name, prm, ct)


def test_pack_cache_refreshes_in_place_and_follows_raw_pointer_updates():
    """ADVICE r1 (high/medium): packed weight copies must follow in-place edits (``_version``), raw-pointer updates
    (``packing.bump_generation``, what FlatAdam.step calls) and keep their addresses so captured graphs stay valid."""
    import torch
    from vognet_pytorch_b200 import packing
    pc = packing.PackCache()
    p = torch.nn.Parameter(torch.ones(3))
    v = pc.get('w', (p,), lambda: p.detach() * 2)
    addr = v.data_ptr()
    assert pc.get('w', (p,), lambda: 1 / 0) is v                       # fresh: build is not called
    with torch.no_grad():
        p.add_(1)
    v2 = pc.get('w', (p,), lambda: p.detach() * 2)
    assert v2.data_ptr() == addr and v2.tolist() == [4.0, 4.0, 4.0]
    p.data[0] = 5                                                      # no version bump through .data
    sig = packing.params_signature([p])
    packing.bump_generation()
    assert packing.params_signature([p]) != sig
    assert pc.refresh() == 1 and v2.tolist() == [10.0, 4.0, 4.0] and pc.relocations == 0
    p.data = torch.zeros(5)                                            # shape change: storage must move
    v3 = pc.get('w', (p,), lambda: p.detach() * 2)
    assert v3.shape == (5,) and pc.relocations == 1


def test_pack_cache_build_into_rederives_in_place_without_temporaries():
    """A training step re-packs every weight after each optimizer update: `build_into` re-derives a stale entry
    straight into the tensors it already owns (same addresses), `refresh()` uses it too, and a `False` return (layout
    changed) falls back to build + relocate."""
    import torch
    from vognet_pytorch_b200 import packing
    pc = packing.PackCache()
    a, b = torch.nn.Parameter(torch.ones(2, 3)), torch.nn.Parameter(torch.full((2, 3), 2.0))
    calls = {'build': 0, 'into': 0}

    def build():
        calls['build'] += 1
        return torch.cat([a.detach(), b.detach()], 0), a.detach().sum(0) + b.detach().sum(0)

    def into(dst):
        calls['into'] += 1
        if dst[0].shape[0] != a.shape[0] + b.shape[0]:
            return False
        dst[0][:a.shape[0]].copy_(a); dst[0][a.shape[0]:].copy_(b)
        dst[1].copy_(a.detach().sum(0) + b.detach().sum(0))
    v = pc.get('k', (a, b), build, into)
    addr = v[0].data_ptr()
    assert calls == {'build': 1, 'into': 0}
    with torch.no_grad():
        a.mul_(3)
    v2 = pc.get('k', (a, b), build, into)
    assert v2 is v and v[0].data_ptr() == addr and calls == {'build': 1, 'into': 1}
    assert v[0][:2].eq(3).all() and v[0][2:].eq(2).all() and v[1].tolist() == [10.0, 10.0, 10.0]
    packing.bump_generation()
    with torch.no_grad():
        b.data.fill_(1.0)
    assert pc.refresh() == 1 and calls['into'] == 2 and v[1].tolist() == [8.0, 8.0, 8.0]
    a.data = torch.ones(4, 3)                                          # layout change: into declines, build relocates
    v3 = pc.get('k', (a, b), build, into)
    assert v3[0].shape == (6, 3) and calls['build'] == 2 and pc.relocations == 1


def test_packed_batch_layout_round_trips_and_exposes_the_graph_prefix():
    """runtime.PackedLayout / pack_host_batch: one flat buffer per batch (a single host->device copy), tensors as views,
    the model's graph inputs as a contiguous prefix."""
    import torch
    from vognet_pytorch_b200 import runtime, synth
    w, batch = synth.workload('cpu_ref')
    first = ('pad_region_feature', 'seg_feature_for_frms', 'pad_proposals')
    pb = runtime.pack_host_batch(batch, first=first, pin=False)
    assert set(pb) == set(batch) and all(torch.equal(pb[k], batch[k]) for k in batch)
    assert all(e[1] % 256 == 0 for e in pb.layout.entries)
    ents, end = pb.layout.prefix(first)
    assert [e[0] for e in ents] == list(first) and end == pb.layout.entries[3][1]
    assert pb.layout.prefix(('pad_proposals',)) is None                     # not a prefix in this order
    # a layout built from the prefix keys alone describes the same bytes (what the captured forward allocates)
    sub = runtime.PackedLayout({k: batch[k] for k in first}, first=first)
    assert sub.prefix(first) == (ents, end) and sub.nbytes == end
    # views alias the flat buffer
    pb['pad_proposals'].zero_()
    o, n = pb.layout.entries[2][1], pb.layout.entries[2][2]
    assert int(pb.flat[o:o + n].sum()) == 0
    with pytest.raises(ValueError):
        pb.layout.pack_into(pb.flat, {**batch, 'pad_proposals': batch['pad_proposals'][:, :1]})


def test_prediction_fetcher_recognises_views_of_one_allocation():
    """runtime.PredictionFetcher moves tensors that are contiguous views of ONE allocation (what ops.select_fwd returns)
    with a single copy; anything else - different storages, a storage that holds more than the tensors - one copy each."""
    import torch
    from vognet_pytorch_b200.runtime import PredictionFetcher
    flat = torch.zeros(8 * 3 + 4 * 6 + 4 * 2, dtype=torch.uint8)
    ix = flat[:24].view(torch.int64)
    bx = flat[24:48].view(torch.float32).view(2, 3)
    sc = flat[48:].view(torch.float32)
    assert PredictionFetcher._shared((bx, sc, ix))
    assert not PredictionFetcher._shared((bx, sc))                       # the allocation holds something else too
    assert not PredictionFetcher._shared((bx, torch.zeros(2)))           # different storages
    big = torch.zeros(4, 4)
    assert not PredictionFetcher._shared((big[:, :2], big[:, 2:]))       # non-contiguous views
