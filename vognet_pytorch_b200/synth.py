"""Deterministic synthetic workloads for the VOGNet forward fusion path.

There is no dataset and no checkpoint on the GPU box, so every test, the smoke
run and ``bench.py`` build their weights and batches here.  Everything is drawn
from numpy's PCG64 bit stream through ``Generator.random(dtype=float32)`` (an
exact integer -> float conversion) followed by plain IEEE float32 arithmetic, so
the same (name, seed) gives bit-identical tensors in this container (where the
golden vectors are generated from the real reference) and on the GPU box.

Shapes follow the batch dict produced by the reference collator
(code/dat_loader_simple.py:1518-1544) and read by the forward
(code/mdl_vog.py:74-89,112,137,296,309,497-506,624; code/mdl_conc_single.py:25,42,77,81,119):

  srl_arg_words_ind      [B,1,nsrl,L]  int64   word ids of every SRL argument, padded to L
  srl_arg_word_mask      [B,1,L]       int64   sentence position -> flat (arg*L+word) index, -1 = pad
  srl_tag_word_ind       [B,1,L]       int64
  srl_arg_word_mask_len  [B,1]         int64   sentence length
  srl_arg_words_capture  [B,1,nsrl,2]  int64   first / last sentence position of each argument
  srl_arg_inds_msk       [B,1,nsrl]    int64   1 = argument slot is populated
  pad_region_feature     [B,P,2048]    float32
  seg_feature_for_frms   [B,ncmp*nfrm,3072] float32
  pad_proposals          [B,P,7]       float32 (x1,y1,x2,y2,frame,class,score  dat_loader_simple.py:191-201)
  new_srl_idxs           [B,ncmp]      int64   (only its size is read)
  num_cmp_msk            [B,ncmp]      int64

Proposal rows are ordered [frame][vid][prop] with x shifted by 720*vid for
``spat`` (dat_loader_simple.py:1067-1103,1151-1153) and [vid][frame][prop] with
the frame id shifted by 10*vid for ``temp`` (dat_loader_simple.py:1231-1252).
"""
import zlib
from collections import OrderedDict

import numpy as np
import torch


class AttrDict(dict):
    """Attribute-access dict standing in for yacs CfgNode / munch.Munch."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    return d


# The cfg keys the forward path reads (configs/anet_srl_cfg.yml:16-24,57-81,98-104),
# with use_rel=True as in every published command (README.md:53, EXPTS.md:14-15).
def default_cfg(conc_type='spat', n_layers=1, n_heads=3, use_rel=True, obj_one_frm=False):
    return to_attr({
        'ds': {'conc_type': conc_type, 'resized_width': 720, 'resized_height': 405,
               'num_sampled_frm': 10, 't_attn_size': 480, 'max_seq_length': 20, 'max_gt_box': 100},
        'mdl': {
            'name': 'vog', 'seg_feat_dim': 3072, 'prop_feat_dim': 2048,
            'input_encoding_size': 512,
            'rnn': {'rnn_size': 1024, 'num_layers': 2, 'drop_prob_lm': 0.5},
            'vsrl': {'prop_encode_size': 256, 'seg_encode_size': 256, 'lang_encode_size': 256},
            'obj_tx': {'use_ddp': False, 'to_use': True, 'n_layers': n_layers, 'n_heads': n_heads,
                       'attn_drop': 0.2, 'use_rel': use_rel, 'one_frm': obj_one_frm},
            'mul_tx': {'use_ddp': False, 'to_use': True, 'n_layers': n_layers, 'n_heads': n_heads,
                       'attn_drop': 0.2, 'use_rel': use_rel, 'one_frm': True, 'cross_frm': False},
        },
        'misc': {'srl_arg_length': 5, 'box_per_srl_arg': 4},
        'loss': {'loss_lambda': 1},
    })


def default_comm(nppf, vocab_size=1000):
    return AttrDict(vocab_size=vocab_size, detect_size=10, itod={}, wtoi={'UNK': 0},
                    num_prop_per_frm=nppf)


# BASELINE.json configs.  'B' is the per-GPU batch.
WORKLOADS = OrderedDict([
    ('cpu_ref',   dict(conc_type='spat', B=1, ncmp=1, nppf=5,   nvalid=4, compute='fp32')),
    ('spat_gt5',  dict(conc_type='spat', B=4, ncmp=4, nppf=5,   nvalid=None, compute='fp32')),
    ('spat_p100', dict(conc_type='spat', B=4, ncmp=4, nppf=100, nvalid=None, compute='bf16')),
    ('temp_gt5',  dict(conc_type='temp', B=4, ncmp=4, nppf=5,   nvalid=None, compute='fp32')),
    ('temp_p100', dict(conc_type='temp', B=4, ncmp=4, nppf=100, nvalid=None, compute='bf16')),
])

# conc_type 'sep' (code/mdl_conc_sep.py): every contrastive video is scored on its own,
# batch tensors carry an extra [ncmp] axis and every video has the sentence appended (append_everywhere,
# code/dat_loader_simple.py:942-945).  Not BASELINE configs: parity cases for the widened path.
WORKLOADS_SEP = OrderedDict([
    ('sep_gt5',  dict(conc_type='sep', B=2, ncmp=4, nppf=5,   nvalid=None, compute='fp32')),
    ('sep_p100', dict(conc_type='sep', B=1, ncmp=2, nppf=100, nvalid=None, compute='bf16')),
])

NFRM0 = 10     # ds.num_sampled_frm
NSRL = 5       # misc.srl_arg_length
SEQ_L = 20     # ds.max_seq_length
WORDS_PER_ARG = 4


def _rng(name, seed):
    return np.random.Generator(np.random.PCG64([zlib.crc32(name.encode()), seed]))


def _uniform(name, seed, shape, lo, hi):
    u = _rng(name, seed).random(size=shape, dtype=np.float32)
    return (u * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------
def param_shapes(nppf_unused=None, vocab_size=1000, n_layers_obj=1, n_layers_mul=1, n_heads=3):
    """state_dict names/shapes of VOG_{SPAT,TEMP} (checkpoint contract)."""
    S = OrderedDict()
    S['lstm_encoder.embed_tokens.weight'] = (vocab_size + 1, 512)
    for layer, in_dim in ((0, 512), (1, 2048)):
        for sfx in ('', '_reverse'):
            S[f'lstm_encoder.lstm.weight_ih_l{layer}{sfx}'] = (4096, in_dim)
            S[f'lstm_encoder.lstm.weight_hh_l{layer}{sfx}'] = (4096, 1024)
            S[f'lstm_encoder.lstm.bias_ih_l{layer}{sfx}'] = (4096,)
            S[f'lstm_encoder.lstm.bias_hh_l{layer}{sfx}'] = (4096,)
    S['lstm_out_feat_proj.0.weight'] = (256, 2048)
    S['lstm_out_feat_proj.0.bias'] = (256,)
    S['srl_arg_words_out_enc.0.weight'] = (256, 512)
    S['srl_arg_words_out_enc.0.bias'] = (256,)
    S['srl_simple_lin.0.weight'] = (256, 768)
    S['srl_simple_lin.0.bias'] = (256,)
    S['prop_encoder.0.weight'] = (256, 2048)
    S['prop_encoder.0.bias'] = (256,)
    S['seg_encoder.0.weight'] = (256, 3072)
    S['seg_encoder.0.bias'] = (256,)
    S['seg_verb_classf.0.weight'] = (256, 512)
    S['seg_verb_classf.0.bias'] = (256,)
    S['seg_verb_classf.2.weight'] = (1, 256)
    S['seg_verb_classf.2.bias'] = (1,)

    def tx(prefix, d, n_layers):
        f = d // 2
        for l in range(n_layers):
            p = f'{prefix}.encoder.layers.{l}'
            for w in ('wq', 'wk', 'wv', 'wo'):
                S[f'{p}.selfattn.layer.{w}.weight'] = (d, d)
            S[f'{p}.selfattn.layernorm.weight'] = (d,)
            S[f'{p}.selfattn.layernorm.bias'] = (d,)
            S[f'{p}.feedforward.layer.linear1.weight'] = (f, d)
            S[f'{p}.feedforward.layer.linear1.bias'] = (f,)
            S[f'{p}.feedforward.layer.linear2.weight'] = (d, f)
            S[f'{p}.feedforward.layer.linear2.bias'] = (d,)
            S[f'{p}.feedforward.layernorm.weight'] = (d,)
            S[f'{p}.feedforward.layernorm.bias'] = (d,)

    tx('obj_txf', 512, n_layers_obj)
    S['pe_obj_sub_enc.0.weight'] = (n_heads, 5)
    S['pe_obj_sub_enc.0.bias'] = (n_heads,)
    for nm in ('lin2', 'lin_tmp'):
        S[f'{nm}.0.weight'] = (256, 768)
        S[f'{nm}.0.bias'] = (256,)
        S[f'{nm}.2.weight'] = (1, 256)
        S[f'{nm}.2.bias'] = (1,)
    tx('mult_txf', 768, n_layers_mul)
    S['pe_mul_sub_enc.0.weight'] = (n_heads, 5)
    S['pe_mul_sub_enc.0.bias'] = (n_heads,)
    return S


def make_state_dict(seed=0, **kw):
    """Synthetic checkpoint.  torch-default-like uniform(+-1/sqrt(fan_in)) init, sharpened where
    the default init would make the forward insensitive to the very things the parity tests have
    to catch: wq/wk x6 (obj) / x4 (mul) so attention logits get O(1) spread instead of ~0.05, pe
    encoders x8 (obj) / x32 (mul) so the relative-position bias visibly changes the softmax, lin2.2 x4 (top-2 score gaps widen so argmax
    selection is meaningful), LayerNorm affine != identity."""
    sd = OrderedDict()
    for name, shape in param_shapes(**kw).items():
        if 'layernorm.weight' in name:
            w = 1.0 + _uniform(name, seed, shape, -0.1, 0.1)
        elif 'layernorm.bias' in name:
            w = _uniform(name, seed, shape, -0.1, 0.1)
        elif name.endswith('embed_tokens.weight'):
            w = _uniform(name, seed, shape, -1.7320508, 1.7320508)
            w[-1] = 0.0  # padding_idx row (utils/mdl_srl_utils.py:93-96)
        elif 'lstm.' in name:
            w = _uniform(name, seed, shape, -1.0 / 32.0, 1.0 / 32.0)  # 1/sqrt(hidden=1024)
        elif name.endswith('.bias'):
            w = _uniform(name, seed, shape, -0.05, 0.05)
        else:
            a = 1.0 / np.sqrt(shape[1])
            w = _uniform(name, seed, shape, -a, a)
            if '.wq.' in name or '.wk.' in name:
                w = w * np.float32(6.0 if name.startswith('obj_txf') else 4.0)
            if name.startswith('pe_'):
                w = w * np.float32(8.0 if name.startswith('pe_obj') else 32.0)
            if name == 'lin2.2.weight':
                w = w * np.float32(4.0)
        sd[name] = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
    return sd


# ---------------------------------------------------------------------------------------------
# inputs
# ---------------------------------------------------------------------------------------------
def make_batch(conc_type='spat', B=4, ncmp=4, nppf=5, nvalid=None, seed=1, vocab_size=1000,
               **_unused):
    """One synthetic batch dict (CPU tensors)."""
    nfrm = NFRM0
    P = ncmp * nfrm * nppf
    r = _rng(f'batch/{conc_type}/{B}/{ncmp}/{nppf}', seed)

    # ---- language side -------------------------------------------------------------------
    if nvalid is None:
        nval = r.integers(2, NSRL + 1, size=B)           # 2..5 populated argument slots
    else:
        nval = np.full(B, nvalid)
    words = r.integers(0, vocab_size, size=(B, 1, NSRL, SEQ_L)).astype(np.int64)
    word_mask = np.full((B, 1, SEQ_L), -1, np.int64)
    capture = np.zeros((B, 1, NSRL, 2), np.int64)
    inds_msk = np.zeros((B, 1, NSRL), np.int64)
    lens = np.zeros((B, 1), np.int64)
    for b in range(B):
        pos = 0
        for s in range(int(nval[b])):
            nw = WORDS_PER_ARG if s % 2 == 0 else WORDS_PER_ARG - 1   # ragged argument lengths
            for w in range(nw):
                word_mask[b, 0, pos + w] = s * SEQ_L + w
            capture[b, 0, s] = (pos, pos + nw - 1)
            inds_msk[b, 0, s] = 1
            pos += nw
        lens[b, 0] = pos
    tags = r.integers(0, 8, size=(B, 1, SEQ_L)).astype(np.int64)

    # ---- visual side ---------------------------------------------------------------------
    feat = np.abs(r.standard_normal(size=(B, P, 2048), dtype=np.float32))
    seg = np.abs(r.standard_normal(size=(B, ncmp * nfrm, 3072), dtype=np.float32))
    x1 = r.random(size=(B, P), dtype=np.float32) * np.float32(600.0)
    y1 = r.random(size=(B, P), dtype=np.float32) * np.float32(300.0)
    w_ = r.random(size=(B, P), dtype=np.float32) * np.float32(100.0) + np.float32(20.0)
    h_ = r.random(size=(B, P), dtype=np.float32) * np.float32(80.0) + np.float32(20.0)
    cls = r.integers(0, 10, size=(B, P)).astype(np.float32)
    score = r.random(size=(B, P), dtype=np.float32)
    idx = np.arange(P)
    if conc_type == 'spat':            # rows [frame][vid][prop]
        frm = idx // (ncmp * nppf)
        vid = (idx // nppf) % ncmp
        xoff = (720.0 * vid).astype(np.float32)
        frame_id = frm.astype(np.float32)
    else:                              # temp: rows [vid][frame][prop]
        vid = idx // (nfrm * nppf)
        frm = (idx // nppf) % nfrm
        xoff = np.zeros(P, np.float32)
        frame_id = (frm + nfrm * vid).astype(np.float32)
    x2 = np.minimum(x1 + w_, np.float32(719.0))
    y2 = np.minimum(y1 + h_, np.float32(404.0))
    props = np.stack([x1 + xoff, y1, x2 + xoff, y2,
                      np.broadcast_to(frame_id, (B, P)), cls, score], axis=-1).astype(np.float32)

    t = torch.from_numpy
    return {
        'srl_arg_words_ind': t(words),
        'srl_arg_word_mask': t(word_mask),
        'srl_tag_word_ind': t(tags),
        'srl_arg_word_mask_len': t(lens),
        'srl_arg_words_capture': t(capture),
        'srl_arg_inds_msk': t(inds_msk),
        'pad_region_feature': t(np.ascontiguousarray(feat)),
        'seg_feature_for_frms': t(np.ascontiguousarray(seg)),
        'pad_proposals': t(np.ascontiguousarray(props)),
        'new_srl_idxs': torch.zeros(B, ncmp, dtype=torch.int64),
        'num_cmp_msk': torch.ones(B, ncmp, dtype=torch.int64),
    }


MAX_GT_BOX = 100       # ds.max_gt_box
BOX_PER_SRL = 4        # misc.box_per_srl_arg


def make_loss_inputs(batch, conc_type='spat', ncmp=4, nppf=5, seed=1, num_box=12, **_unused):
    """The extra batch keys the reference loss reads (code/mdl_conc_single.py:191-311, layouts from
    code/dat_loader_simple.py:405-456,703-779,1170-1200): ground-truth boxes, frame / padding masks, the gt-box
    indices of every SRL argument, the target video.  Drawn from their own RNG stream, so the forward batch of
    ``make_batch`` stays bit-identical.  Ground-truth boxes are jittered copies of proposals of the target
    video placed on the NEXT frame: ``pad_frm_mask`` is 1 where the frames of a proposal and a gt box DIFFER
    (dat_loader_simple.py:237-255) and the reference multiplies the overlaps by it (utils/box_utils.py:108-110),
    so this is what produces positive targets."""
    props = batch['pad_proposals'].numpy()
    B, P, _ = props.shape
    nfrm = NFRM0
    r = _rng(f'loss/{conc_type}/{B}/{ncmp}/{nppf}', seed)
    idx = np.arange(P)
    vid = (idx // nppf) % ncmp if conc_type == 'spat' else idx // (nfrm * nppf)
    gt = np.zeros((B, MAX_GT_BOX, 5), np.float32)
    target_cmp = r.integers(0, ncmp, size=B).astype(np.int64)
    srl_boxes = np.zeros((B, 1, NSRL, BOX_PER_SRL), np.int64)
    srl_lens = np.zeros((B, 1, NSRL, BOX_PER_SRL), np.int64)
    arg_boxes_mask = np.zeros((B, 1, NSRL), np.int64)
    inds = batch['srl_arg_inds_msk'].numpy()
    frame_col = props[..., 4]
    fmin = frame_col.min()
    for b in range(B):
        cand = np.nonzero(vid == target_cmp[b])[0]
        pick = r.choice(cand, size=num_box, replace=len(cand) < num_box)
        jit = r.integers(-6, 7, size=(num_box, 4)).astype(np.float32)
        gt[b, :num_box, :4] = props[b, pick, :4] + jit
        # next frame of the same video (wraps inside the video's own frame range)
        f = frame_col[b, pick]
        base = np.floor((f - fmin) / nfrm) * nfrm + fmin if conc_type == 'temp' else fmin
        gt[b, :num_box, 4] = base + np.mod(f - base + 1, nfrm)
        for s_ in range(NSRL):
            if inds[b, 0, s_] and s_ % 3 != 2:                 # some populated arguments are not groundable
                n = int(r.integers(1, BOX_PER_SRL + 1))
                srl_boxes[b, 0, s_, :n] = r.integers(0, num_box, size=n)
                srl_lens[b, 0, s_, :n] = 1
                arg_boxes_mask[b, 0, s_] = 1
    frm_mask = np.ones((B, P, MAX_GT_BOX), np.uint8)
    frm_mask[:, :, :num_box] = (frame_col[:, :, None] != gt[:, None, :num_box, 4]).astype(np.uint8)
    pnt_mask = (r.random(size=(B, P)) < 0.02).astype(np.uint8)   # a few padded proposal slots
    t = torch.from_numpy
    return {'pad_gt_bboxs': t(gt), 'pad_frm_mask': t(frm_mask), 'pad_pnt_mask': t(pnt_mask),
            'srl_boxes': t(srl_boxes), 'srl_boxes_lens': t(srl_lens), 'srl_arg_boxes_mask': t(arg_boxes_mask),
            'target_cmp': t(target_cmp)}


def make_loss_inputs_sep(batch, ncmp=4, nppf=5, seed=1, num_box=12, **_unused):
    """Loss inputs of LossB_SEP (code/mdl_conc_sep.py:236-447): the single-video loss inputs of every (query, video)
    pair with the [ncmp] axis unfolded (gt boxes, masks and SRL box indices per video), the target video of every
    query, and the verb targets ``verb_cmp`` [B,ncmp] / ``verb_cross_cmp_msk`` [B,ncmp,ncmp]."""
    B = batch['pad_proposals'].shape[0]
    Bq = B * ncmp
    flat = {'pad_proposals': batch['pad_proposals'].reshape(Bq, *batch['pad_proposals'].shape[2:]),
            'srl_arg_inds_msk': batch['srl_arg_inds_msk'].reshape(Bq, 1, -1)}
    one = make_loss_inputs(flat, conc_type='temp', ncmp=1, nppf=nppf, seed=seed, num_box=num_box)
    out = {}
    for k, v in one.items():
        if k == 'target_cmp':
            continue
        out[k] = (v.reshape(B, ncmp, *v.shape[2:]) if k in ('srl_boxes', 'srl_boxes_lens', 'srl_arg_boxes_mask')
                  else v.reshape(B, ncmp, *v.shape[1:])).contiguous()
    r = _rng(f'loss_sep/{B}/{ncmp}/{nppf}', seed)
    out['target_cmp'] = torch.from_numpy(r.integers(0, ncmp, size=B).astype(np.int64))
    out['verb_cmp'] = torch.from_numpy(r.integers(0, 2, size=(B, ncmp)).astype(np.int64))
    vcc = r.integers(0, 2, size=(B, ncmp, ncmp)).astype(np.int64)
    vcc[:, 0, :] = 0                                    # one video per query without a verb target
    vcc[:, 1:, 0] = 1
    out['verb_cross_cmp_msk'] = torch.from_numpy(vcc)
    return out


def make_batch_sep(B=2, ncmp=4, nppf=5, nvalid=None, seed=1, vocab_size=1000, **_unused):
    """SEP batch (code/mdl_conc_sep.py:131-217 reads these shapes): the B*ncmp (query, video) pairs of a
    single-video batch with the [ncmp] axis unfolded -

      srl_arg_words_ind [B,ncmp,nsrl,L]   srl_arg_word_mask [B,ncmp,L]   srl_arg_word_mask_len [B,ncmp]
      srl_arg_words_capture [B,ncmp,nsrl,2]   srl_arg_inds_msk [B,ncmp,nsrl]   verb_ind_in_srl [B,ncmp]
      pad_region_feature [B,ncmp,nfrm*nppf,2048]   seg_feature_for_frms [B,ncmp,nfrm,3072]
      pad_proposals [B,ncmp,nfrm*nppf,7]   new_srl_idxs / num_cmp_msk [B,ncmp]

    Every video keeps its own frame ids 0..nfrm-1 and un-shifted boxes.  The last video of odd queries is masked
    out (num_cmp_msk = 0) to exercise the mask path."""
    one = make_batch(conc_type='temp', B=B * ncmp, ncmp=1, nppf=nppf, nvalid=nvalid, seed=seed,
                     vocab_size=vocab_size)
    out = {}
    for k, v in one.items():
        if k in ('new_srl_idxs', 'num_cmp_msk'):
            continue
        if k in ('pad_region_feature', 'seg_feature_for_frms', 'pad_proposals'):
            out[k] = v.reshape(B, ncmp, *v.shape[1:]).contiguous()
        else:                                           # language tensors [B*ncmp, 1, ...] -> [B, ncmp, ...]
            out[k] = v.reshape(B, ncmp, *v.shape[2:]).contiguous()
    r = _rng(f'batch_sep/{B}/{ncmp}/{nppf}', seed)
    nval = out['srl_arg_inds_msk'].sum(-1).numpy()     # populated argument slots of every (query, video)
    verb = (r.integers(0, 1 << 30, size=(B, ncmp)) % nval).astype(np.int64)
    out['verb_ind_in_srl'] = torch.from_numpy(verb)
    out['new_srl_idxs'] = torch.zeros(B, ncmp, dtype=torch.int64)
    msk = torch.ones(B, ncmp, dtype=torch.int64)
    if ncmp > 1:
        msk[1::2, -1] = 0
    out['num_cmp_msk'] = msk
    return out


def clone_batch(batch, device=None):
    """The reference forward mutates ``srl_arg_word_mask`` in place (code/mdl_vog.py:80-82)."""
    return {k: (v.clone() if device is None else v.to(device, copy=True)) for k, v in batch.items()}


def workload(name, seed=1):
    if name in WORKLOADS_SEP:
        w = dict(WORKLOADS_SEP[name])
        return w, make_batch_sep(seed=seed, **w)
    w = dict(WORKLOADS[name])
    return w, make_batch(seed=seed, **w)


# ---------------------------------------------------------------------------------------------
# operator-level cases (RelTransformer / Transformer called directly)
# ---------------------------------------------------------------------------------------------
def make_operator_state_dict(d, n_layers, seed=3):
    """state_dict of a bare RelTransformer/Transformer (names 'encoder.layers.{l}....')."""
    sd = OrderedDict()
    f = d // 2
    for l in range(n_layers):
        p = f'encoder.layers.{l}'
        a = 1.0 / np.sqrt(d)
        for w in ('wq', 'wk', 'wv', 'wo'):
            m = _uniform(f'op{d}/{p}.{w}', seed, (d, d), -a, a)
            if w in ('wq', 'wk'):
                m = m * np.float32(5.0)
            sd[f'{p}.selfattn.layer.{w}.weight'] = m
        sd[f'{p}.selfattn.layernorm.weight'] = 1.0 + _uniform(f'op{d}/{p}.ln1w', seed, (d,), -0.1, 0.1)
        sd[f'{p}.selfattn.layernorm.bias'] = _uniform(f'op{d}/{p}.ln1b', seed, (d,), -0.1, 0.1)
        sd[f'{p}.feedforward.layer.linear1.weight'] = _uniform(f'op{d}/{p}.l1w', seed, (f, d), -a, a)
        sd[f'{p}.feedforward.layer.linear1.bias'] = _uniform(f'op{d}/{p}.l1b', seed, (f,), -0.05, 0.05)
        a2 = 1.0 / np.sqrt(f)
        sd[f'{p}.feedforward.layer.linear2.weight'] = _uniform(f'op{d}/{p}.l2w', seed, (d, f), -a2, a2)
        sd[f'{p}.feedforward.layer.linear2.bias'] = _uniform(f'op{d}/{p}.l2b', seed, (d,), -0.05, 0.05)
        sd[f'{p}.feedforward.layernorm.weight'] = 1.0 + _uniform(f'op{d}/{p}.ln2w', seed, (d,), -0.1, 0.1)
        sd[f'{p}.feedforward.layernorm.bias'] = _uniform(f'op{d}/{p}.ln2b', seed, (d,), -0.1, 0.1)
    return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)))
                       for k, v in sd.items())


def make_operator_inputs(d, n_heads, Bt, N, seed=3):
    """x [Bt,N,d] ~ U(-1,1); dense bias x_pe [Bt,N,N,H] ~ U(0,8) with ~half the entries clamped
    to 0 (it is a ReLU output in the model)."""
    x = _uniform(f'opx/{d}/{Bt}/{N}', seed, (Bt, N, d), -1.0, 1.0)
    pe = np.maximum(_uniform(f'oppe/{n_heads}/{Bt}/{N}', seed, (Bt, N, N, n_heads), -8.0, 8.0),
                    np.float32(0.0))
    return torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(pe))
