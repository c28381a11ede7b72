"""TEST INFRASTRUCTURE ONLY - imports the UNMODIFIED reference from /root/reference.

Root of the reference tree: ``$VOG_REFERENCE_ROOT``, else ``/root/reference`` (build container), else
``baseline/_ref`` - the git-ignored copy of the reference's ``code/``, ``utils/`` and ``configs/`` that
``__graft_entry__.build()`` installs so that the unmodified reference travels to the GPU box (the
reference has no setup.py / pyproject, so ``pip install --target baseline/_ref`` has nothing to install;
the files are copied verbatim instead and never enter the git history).  Used by ``oracle/make_golden.py``
to produce the committed golden vectors under ``tests/golden/``, by ``tests/test_oracle_vs_reference.py``
(skipped when no reference tree is present) to pin the restatement in ``oracle/vog_oracle.py`` against the
real code, and by ``bench.py`` for its two baseline legs (``--impl reference`` on the host cores and
``gpu_eager_baseline`` on cuda:0).  Nothing in the product path may import this module.

The reference needs two third-party modules that are not installed and carry no arithmetic on
this path: ``munch.Munch`` (attribute dict; pinned munch==2.5.0,
conda_env_vog.yml:159) and ``fairseq.utils`` (imported by utils/mdl_srl_utils.py:8, only called
under left_pad=True which code/mdl_vog.py:172 never sets; pinned fairseq==0.8.0).  The evaluator
module additionally pulls yacs / fire / fastprogress at import time; all are stubbed.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED_ROOT = os.path.join(_REPO, 'baseline', '_ref')


def _pick_root():
    for r in (os.environ.get('VOG_REFERENCE_ROOT'), '/root/reference', INSTALLED_ROOT):
        if r and os.path.isfile(os.path.join(r, 'code', 'mdl_vog.py')):
            return r
    return '/root/reference'


REF_ROOT = _pick_root()


def install_reference(src='/root/reference', dst=INSTALLED_ROOT):
    """Copy the reference's python sources + config verbatim into baseline/_ref (git-ignored, travels with the
    gpurun snapshot).  No-op when the source tree is absent.  -> number of files copied."""
    import shutil
    if not os.path.isfile(os.path.join(src, 'code', 'mdl_vog.py')):
        return 0
    n = 0
    for sub, pat in (('code', '.py'), ('utils', '.py'), ('configs', '.yml')):
        os.makedirs(os.path.join(dst, sub), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(src, sub))):
            if f.endswith(pat):
                shutil.copyfile(os.path.join(src, sub, f), os.path.join(dst, sub, f))
                n += 1
    return n


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, 'code', 'mdl_vog.py'))


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _install_stubs():
    if 'munch' not in sys.modules:
        sys.modules['munch'] = types.SimpleNamespace(Munch=Munch)
    if 'fairseq' not in sys.modules:
        m = types.ModuleType('fairseq')
        m.utils = None
        sys.modules['fairseq'] = m

    def stub(name, **attrs):
        if name in sys.modules:
            return
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m

    stub('fire', Fire=lambda *a, **k: None)
    stub('yacs')
    if isinstance(sys.modules.get('yacs'), types.ModuleType) and not hasattr(sys.modules['yacs'], 'config'):
        cfgm = types.ModuleType('yacs.config')
        cfgm.CfgNode = Munch
        sys.modules['yacs'].config = cfgm
        sys.modules['yacs.config'] = cfgm
    stub('fastprogress', progress_bar=lambda x, **k: x, master_bar=lambda x, **k: x)
    if 'fastprogress.fastprogress' not in sys.modules:
        fp = types.ModuleType('fastprogress.fastprogress')
        fp.progress_bar = lambda x, **k: x
        fp.master_bar = lambda x, **k: x
        fp.MasterBar = object
        fp.ProgressBar = object
        fp.format_time = lambda t: str(t)
        sys.modules['fastprogress.fastprogress'] = fp
        sys.modules['fastprogress'].fastprogress = fp
    for p in (os.path.join(REF_ROOT, 'code'), os.path.join(REF_ROOT, 'utils')):
        if p not in sys.path:          # what code/_init_stuff.py:38-39 does
            sys.path.append(p)


def to_munch(d):
    if isinstance(d, dict):
        return Munch({k: to_munch(v) for k, v in d.items()})
    return d


def reference_cfg(conc_type='spat', n_layers=1, n_heads=3, use_rel=True, obj_one_frm=False):
    import yaml
    cfg = to_munch(yaml.safe_load(open(os.path.join(REF_ROOT, 'configs', 'anet_srl_cfg.yml'))))
    cfg.ds.conc_type = conc_type
    for tx in (cfg.mdl.obj_tx, cfg.mdl.mul_tx):
        tx.use_rel = use_rel
        tx.n_layers = n_layers
        tx.n_heads = n_heads
    cfg.mdl.obj_tx.one_frm = obj_one_frm
    return cfg


def build_reference_model(conc_type, nppf, state_dict, vocab_size=1000, mdl_name='vog', **cfg_kw):
    """Unmodified reference VOG_SPAT / VOG_TEMP / VOG_SEP (or the ImgGrnd_* / VidGrnd_* ablations with
    ``mdl_name='igrnd' | 'vgrnd'``) in eval mode with ``state_dict`` loaded strictly (restricted to the parameters
    the variant owns)."""
    _install_stubs()
    import mdl_vog  # noqa: from /root/reference/code
    cfg = reference_cfg(conc_type, **cfg_kw)
    comm = Munch(vocab_size=vocab_size, detect_size=10, itod={}, wtoi={'UNK': 0},
                 num_prop_per_frm=nppf)
    prefix = {'vog': 'VOG', 'vgrnd': 'VidGrnd', 'igrnd': 'ImgGrnd'}[mdl_name]
    cls = getattr(mdl_vog, f'{prefix}_{conc_type.upper()}')
    cfg.mdl.name = mdl_name
    mdl = cls(cfg, comm)
    own = set(mdl.state_dict())
    mdl.load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=True)
    return mdl.eval()


def build_reference_evaluator(conc_type, nppf, ncmp):
    """EvaluatorSPAT/TEMP without its dataset-reading ctor (code/eval_fn_corr.py:54-67)."""
    _install_stubs()
    import eval_vsrl_corr  # noqa
    cls = {'spat': eval_vsrl_corr.EvaluatorSPAT, 'temp': eval_vsrl_corr.EvaluatorTEMP,
           'sep': eval_vsrl_corr.EvaluatorSEP}[conc_type]
    ev = cls.__new__(cls)
    import torch
    torch.nn.Module.__init__(ev)
    ev.num_sampled_frm = 10
    ev.num_frms = 10
    ev.num_prop_per_frm = nppf
    return ev


def _byte_mask_compat():
    """The reference targets PyTorch 1.1, where ``torch.masked_select`` took uint8 masks
    (code/mdl_conc_single.py:306,410 pass ``boxes_msk.byte()``); torch >= 1.2 wants bool.  Version shim only:
    the mask is reinterpreted, no arithmetic changes."""
    import torch
    if getattr(torch.masked_select, '_vog_compat', False):
        return
    orig = torch.masked_select

    def masked_select(inp, mask, *a, **k):
        return orig(inp, mask.bool() if mask.dtype == torch.uint8 else mask, *a, **k)
    masked_select._vog_compat = True
    torch.masked_select = masked_select


def build_reference_loss(conc_type, nppf):
    """Unmodified LossB_SPAT / LossB_TEMP (code/mdl_conc_single.py:180-433) / LossB_SEP (code/mdl_conc_sep.py:219-447)."""
    _install_stubs()
    _byte_mask_compat()
    import mdl_conc_single  # noqa
    import mdl_conc_sep  # noqa
    cfg = reference_cfg(conc_type)
    comm = Munch(vocab_size=1000, detect_size=10, itod={}, wtoi={'UNK': 0}, num_prop_per_frm=nppf)
    cls = {'spat': mdl_conc_single.LossB_SPAT, 'temp': mdl_conc_single.LossB_TEMP,
           'sep': mdl_conc_sep.LossB_SEP}[conc_type]
    return cls(cfg, comm)


def reference_transformers():
    _install_stubs()
    import transformer_code  # noqa
    return transformer_code


def reference_item_closures(conc_type, nfrm, nppf):
    """The nested helpers ``reshuffle_boxes`` / ``process_props`` of ``verb_item_getter_SPAT`` / ``_TEMP``
    (code/dat_loader_simple.py:1067-1103 / :1231-1252), compiled from the UNMODIFIED source text: they are closures
    of a dataset method whose class needs the 530 GB dataset to construct, so their definitions are lifted out of
    the parsed file and executed with a stand-in ``self`` carrying the two attributes they read."""
    import ast
    import torch
    path = os.path.join(REF_ROOT, 'code', 'dat_loader_simple.py')
    tree = ast.parse(open(path).read())
    getter = 'verb_item_getter_' + conc_type.upper()
    meth = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == getter)
    defs = [n for n in meth.body if isinstance(n, ast.FunctionDef) and n.name in ('reshuffle_boxes', 'process_props')]
    ns = {'torch': torch, 'self': types.SimpleNamespace(num_frms=nfrm, num_prop_per_frm=nppf)}
    exec(compile(ast.Module(body=defs, type_ignores=[]), path, 'exec'), ns)
    _install_stubs()
    import mdl_srl_utils  # noqa: utils/mdl_srl_utils.py:11-17 combine_first_ax
    ns['combine_first_ax'] = mdl_srl_utils.combine_first_ax
    return ns


def reference_concat_videos(batch, conc_type, nfrm, nppf):
    """What the reference's item getters do to the per-video visual tensors of every sample
    (code/dat_loader_simple.py:1147-1153,1196-1207 SPAT; :1290-1292 + stacking TEMP), stacked over the batch."""
    import torch
    ns = reference_item_closures(conc_type, nfrm, nppf)
    feats, segs, props = [], [], []
    for b in range(batch['pad_proposals'].shape[0]):
        p, f, s = batch['pad_proposals'][b], batch['pad_region_feature'][b], batch['seg_feature_for_frms'][b]
        if conc_type == 'spat':
            props.append(ns['process_props'](p, keepdim=False, reshuffle_box=True))
            feats.append(ns['reshuffle_boxes'](f))
            segs.append(ns['combine_first_ax'](s.transpose(0, 1).contiguous(), keepdim=False))
        else:
            props.append(ns['process_props'](p, keepdim=False))
            feats.append(ns['combine_first_ax'](f, keepdim=False))           # temporal stacking: rows stay [vid][frame][prop]
            segs.append(ns['combine_first_ax'](s, keepdim=False))
    return torch.stack(feats), torch.stack(segs), torch.stack(props)
