"""TEST INFRASTRUCTURE ONLY - CPU restatement (torch, fp32) of the reference VOGNet forward
fusion path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this file; the product path never does.

Parity pin: the reference repo ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF,
executed unmodified in the build container by ``oracle/make_golden.py`` (committed together with
the vectors in ``tests/golden/``) and re-checked live by ``tests/test_oracle_vs_reference.py``
whenever /root/reference is present.

The algorithm is restated as stateless functions over a flat ``state_dict`` (the checkpoint
contract) - same operations, same order of floating-point evaluation wherever the order is
observable (bias materialised and added before the 1/sqrt(d_model) scaling, heads split with
uneven ``chunk`` sizes, post-LN residual blocks, eval-mode dropout = identity).  Each function
cites the reference lines it follows.
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# operator level: code/transformer_code.py
# ---------------------------------------------------------------------------------------------
def chunk_sizes(d, n_heads):
    """torch.chunk split sizes (transformer_code.py:66-67,182-183): ceil(d/H) per head, last
    one takes the remainder - d=512,H=3 -> 171,171,170."""
    c = -(-d // n_heads)
    sizes = []
    left = d
    while left > 0:
        sizes.append(min(c, left))
        left -= c
    return sizes


def layer_norm(x, w, b):
    """nn.LayerNorm(d_model), eps=1e-5 (transformer_code.py:28)."""
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def multihead_attention(x, sd, p, n_heads, bias=None):
    """RelMultiHead.forward + RelAttention.forward (transformer_code.py:136-160,176-186) and,
    with bias=None, MultiHead/Attention (:41-50,62-70).

    x [Bt,N,d]; bias [Bt,N,N,H] or None.  softmax((q k^T + bias_h) / sqrt(d_model)) v per head,
    the scale being sqrt(d_model) (NOT sqrt(d_head)): RelAttention is built with d_key=d_model
    (:166,:195) and self.scale = sqrt(d_key) (:132).  No mask is ever applied (causal=False)."""
    d = x.shape[-1]
    q = x @ sd[p + '.wq.weight'].t()
    k = x @ sd[p + '.wk.weight'].t()
    v = x @ sd[p + '.wv.weight'].t()
    scale = math.sqrt(d)
    outs = []
    off = 0
    for h, dh in enumerate(chunk_sizes(d, n_heads)):
        qh, kh, vh = q[..., off:off + dh], k[..., off:off + dh], v[..., off:off + dh]
        off += dh
        dots = torch.matmul(qh, kh.transpose(1, 2))                     # :141
        if bias is not None:
            dots = (dots + bias[..., h]) / scale                          # :149-151
        else:
            dots = dots / scale                                           # :50
        attn = torch.softmax(dots, dim=-1)                                # :153 (dropout off in eval)
        outs.append(torch.matmul(attn, vh))                               # :155
    return torch.cat(outs, -1) @ sd[p + '.wo.weight'].t()               # :184-186


def encoder_layer(x, sd, p, n_heads, bias=None):
    """RelEncoderLayer / EncoderLayer (transformer_code.py:84-96,189-203) with ResidualBlock
    (:21-31) and FeedForward (:73-81): y = LN(x + MHA(x)); out = LN(y + W2 relu(W1 y + b1) + b2)."""
    a = multihead_attention(x, sd, p + '.selfattn.layer', n_heads, bias)
    y = layer_norm(x + a, sd[p + '.selfattn.layernorm.weight'], sd[p + '.selfattn.layernorm.bias'])
    h = F.relu(y @ sd[p + '.feedforward.layer.linear1.weight'].t()
               + sd[p + '.feedforward.layer.linear1.bias'])
    f = h @ sd[p + '.feedforward.layer.linear2.weight'].t() + sd[p + '.feedforward.layer.linear2.bias']
    return layer_norm(y + f, sd[p + '.feedforward.layernorm.weight'],
                      sd[p + '.feedforward.layernorm.bias'])


def transformer(x, sd, prefix, n_heads, bias=None):
    """RelTransformer / Transformer forward: last layer's output (transformer_code.py:227-241,
    :252-254,:271-273)."""
    n_layers = 0
    while f'{prefix}.encoder.layers.{n_layers}.selfattn.layer.wq.weight' in sd:
        n_layers += 1
    for l in range(n_layers):
        x = encoder_layer(x, sd, f'{prefix}.encoder.layers.{l}', n_heads, bias)
    return x


# ---------------------------------------------------------------------------------------------
# relative-position bias: code/mdl_vog.py:456-490 + utils/mdl_srl_utils.py:30-69
# ---------------------------------------------------------------------------------------------
def compute_pe(props5, nsrl, nfrm, nppf, W, b, vid_w=720.0, vid_h=405.0):
    """props5 [B, nfrm*nppf, 5] (x1,y1,x2,y2,frame) -> bias [B*nfrm, nsrl*nppf, nsrl*nppf, H].

    Normalise x/720, y/405, frame/nfrm (mdl_vog.py:459-463), pairwise difference p_i - p_j inside
    each frame group (do_cross 'subtract', mdl_srl_utils.py:55-69), Linear(5,H)+ReLU
    (mdl_vog.py:446-451,480), then tile nsrl x nsrl (:482-488) - token t = s*nppf + p."""
    props = props5.clone()
    props[..., 0] /= vid_w
    props[..., 1] /= vid_h
    props[..., 2] /= vid_w
    props[..., 3] /= vid_h
    props[..., 4] /= nfrm
    B = props.shape[0]
    g = props.view(B * nfrm, nppf, 5)
    diff = g.unsqueeze(2) - g.unsqueeze(1)                  # [.., i, j, :] = p_i - p_j
    pe = F.relu(diff @ W.t() + b)                           # [B*nfrm, nppf, nppf, H]
    H = pe.shape[-1]
    pe = pe.view(B, nfrm, 1, nppf, 1, nppf, H).expand(B, nfrm, nsrl, nppf, nsrl, nppf, H)
    return pe.contiguous().view(B * nfrm, nsrl * nppf, nsrl * nppf, H)


# ---------------------------------------------------------------------------------------------
# language side: code/mdl_vog.py:67-140,250-283 ; utils/mdl_srl_utils.py:114-169
# ---------------------------------------------------------------------------------------------
_LSTM_SKELETON = {}


def _lstm_from_state_dict(sd):
    """nn.LSTM evaluated functionally on the state_dict's own tensors (torch.func.functional_call), so that autograd
    reaches them - the restated path is differentiable end to end for the backward parity checks."""
    w = sd['lstm_encoder.lstm.weight_ih_l0']
    key = (w.shape[1], w.shape[0] // 4)
    if key not in _LSTM_SKELETON:
        _LSTM_SKELETON[key] = torch.nn.LSTM(input_size=key[0], hidden_size=key[1], num_layers=2, dropout=0.0,
                                            bidirectional=True).eval()
    lstm = _LSTM_SKELETON[key]
    params = {k: sd['lstm_encoder.lstm.' + k] for k in lstm.state_dict()}
    return lambda packed: torch.func.functional_call(lstm, params, (packed,))


def language_encode(sd, inp, vocab_size):
    """-> [B, 1, nsrl, 256] argument encodings (the 'language matrix')."""
    words = inp['srl_arg_words_ind']
    B, nv, nsrl, L = words.shape
    flat = words.reshape(B * nv, nsrl * L)
    wm = inp['srl_arg_word_mask'].reshape(B * nv, -1).clone()
    pad = wm == -1
    wm[pad] = 0
    toks = torch.gather(flat, 1, wm)                        # mdl_vog.py:83-85
    toks[pad] = vocab_size                                  # :87
    lens = inp['srl_arg_word_mask_len'].reshape(B * nv)
    toks = toks[:, :int(lens.max())].contiguous()           # :257
    emb = F.embedding(toks, sd['lstm_encoder.embed_tokens.weight'])      # mdl_srl_utils.py:128
    packed = torch.nn.utils.rnn.pack_padded_sequence(emb.transpose(0, 1), lens.tolist(),
                                                     enforce_sorted=False)  # :136-137
    out, _ = _lstm_from_state_dict(sd)(packed)              # zero initial state :145-148
    out, _ = torch.nn.utils.rnn.pad_packed_sequence(out, padding_value=0.0)  # :151-152
    full = out.transpose(0, 1).contiguous()                 # [B, T, 2048]  mdl_vog.py:272-273
    full = F.relu(full @ sd['lstm_out_feat_proj.0.weight'].t() + sd['lstm_out_feat_proj.0.bias'])  # :275
    cap = inp['srl_arg_words_capture'].reshape(B * nv, nsrl, 2)
    D = full.shape[-1]
    st = torch.gather(full, 1, cap[..., 0].unsqueeze(-1).expand(B * nv, nsrl, D))   # :119-120
    en = torch.gather(full, 1, cap[..., 1].unsqueeze(-1).expand(B * nv, nsrl, D))   # :121-122
    enc = torch.cat([st, en], 2).view(B, nv, nsrl, 2 * D)   # :126-128
    enc = F.relu(enc @ sd['srl_arg_words_out_enc.0.weight'].t() + sd['srl_arg_words_out_enc.0.bias'])
    return enc * inp['srl_arg_inds_msk'].unsqueeze(-1).float()            # :137-139


# ---------------------------------------------------------------------------------------------
# whole forward: code/mdl_conc_single.py:68-127 (TEMP) / :130-177 (SPAT)
# ---------------------------------------------------------------------------------------------
def vog_forward(sd, inp, conc_type, nppf, n_heads=3, vocab_size=1000, nfrm0=10,
                use_rel=True, keep=False, mdl_name='vog', obj_one_frm=False):
    """Restated ConcTEMP.forward / ConcSPAT.forward for mdl.name='vog'; ``mdl_name='vgrnd'`` drops the multimodal
    transformer (VidGrnd: conc_encode_simple, code/mdl_vog.py:346-363), ``'igrnd'`` also the object transformer
    (ImgGrnd.simple_obj_interact :285-289).

    returns {'mdl_outs': logits [B,1,nsrl,P], 'mdl_outs_eval': masked sigmoid} (+ intermediates
    when keep=True)."""
    ncmp = inp['new_srl_idxs'].shape[1]
    B, nv, nsrl, _ = inp['srl_arg_words_ind'].shape
    lang = language_encode(sd, inp, vocab_size)                            # [B,1,nsrl,256]

    prop = F.relu(inp['pad_region_feature'] @ sd['prop_encoder.0.weight'].t()
                  + sd['prop_encoder.0.bias'])                             # mdl_vog.py:291-299
    seg = F.relu(inp['seg_feature_for_frms'] @ sd['seg_encoder.0.weight'].t()
                 + sd['seg_encoder.0.bias'])                               # :301-314
    P = prop.shape[1]
    nvf = seg.shape[1]                                                     # ncmp*nfrm0 (frame,vid) slots
    # every proposal of a (frame,vid) slot gets that slot's segment feature
    # (mdl_conc_single.py:50-66,156-174)
    ps = torch.cat([prop.view(B, 1, nvf, nppf, -1),
                    seg.view(B, 1, nvf, 1, -1).expand(B, 1, nvf, nppf, seg.shape[-1])], -1)
    ps = ps.reshape(B, 1, P, -1)                                           # [B,1,P,512]

    # ---- object transformer over ALL proposals of the query (mdl_vog.py:492-523, one_frm False)
    props5 = inp['pad_proposals'][..., :5].clone()
    if obj_one_frm:
        # obj_tx.one_frm (:496-504): one sequence per frame group, the groups of simple_obj_interact_input
        # (mdl_conc_single.py:30-37 TEMP, :137-142 SPAT); frame id divided by the number of groups
        nfo, npo = (nfrm0, ncmp * nppf) if conc_type == 'spat' else (ncmp * nfrm0, nppf)
    else:
        nfo, npo = 1, P          # compute_pe(props, nsrl=1, nfrm=1, nppf=P): frame id divided by 1 (:507-510)
    x_obj = ps.reshape(B * nfo, npo, -1)
    if mdl_name == 'igrnd':
        bias_obj = None
    elif use_rel:
        bias_obj = compute_pe(props5, 1, nfo, npo, sd['pe_obj_sub_enc.0.weight'], sd['pe_obj_sub_enc.0.bias'])
    else:
        bias_obj = None
    y_obj = x_obj if mdl_name == 'igrnd' else transformer(x_obj, sd, 'obj_txf', n_heads, bias_obj)
    del bias_obj
    vis = y_obj.view(B, 1, P, -1)

    # ---- vis/lang concat (mdl_vog.py:316-344): token (s,p) = [vis_p | lang_s]
    conc = torch.cat([vis.view(B, 1, 1, P, -1).expand(B, 1, nsrl, P, vis.shape[-1]),
                      lang.view(B, 1, nsrl, 1, -1).expand(B, 1, nsrl, P, lang.shape[-1])], -1)

    # ---- multimodal transformer, one sequence per frame (mdl_vog.py:681-744, int_pfrm=True)
    if conc_type == 'spat':
        nfrm, nppf2 = nfrm0, ncmp * nppf                                   # mdl_conc_single.py:131-135
    else:
        nfrm, nppf2 = ncmp * nfrm0, nppf                                   # :24-28
    D = conc.shape[-1]
    x_mul = conc.view(B, nsrl, nfrm, nppf2, D).transpose(1, 2).contiguous().view(
        B * nfrm, nsrl * nppf2, D)                                         # mdl_vog.py:693-699
    if mdl_name != 'vog':
        bias_mul = None
    elif use_rel:
        bias_mul = compute_pe(inp['pad_proposals'][..., :5].clone(), nsrl, nfrm, nppf2,
                              sd['pe_mul_sub_enc.0.weight'], sd['pe_mul_sub_enc.0.bias'])  # :710-713
    else:
        bias_mul = None
    # igrnd / vgrnd: lin2 directly on the concatenated tokens (the regroup is a pure permutation around a row-wise map)
    y_mul = x_mul if mdl_name != 'vog' else transformer(x_mul, sd, 'mult_txf', n_heads, bias_mul)
    del bias_mul
    y = y_mul.view(B, nfrm, nsrl, nppf2, D).transpose(1, 2).contiguous().view(B, 1, nsrl, P, D)  # :724-737

    h = F.relu(y @ sd['lin2.0.weight'].t() + sd['lin2.0.bias'])
    logits = (h @ sd['lin2.2.weight'].t() + sd['lin2.2.bias']).squeeze(-1)   # mdl_vog.py:675-677

    # ---- output masks (mdl_conc_single.py:39-48,118-122,144-154)
    cm = inp['num_cmp_msk']
    if conc_type == 'spat':
        cmsk = cm.view(B, 1, 1, 1, ncmp, 1).expand(B, nv, nsrl, nfrm0, ncmp, nppf)
    else:
        cmsk = cm.view(B, 1, 1, ncmp, 1).expand(B, nv, nsrl, ncmp, nfrm0 * nppf)
    cmsk = cmsk.reshape(logits.shape)
    smsk = inp['srl_arg_inds_msk'].unsqueeze(-1).expand(*logits.shape)
    ev = torch.sigmoid(logits) * smsk.float() * cmsk.float()
    out = {'mdl_outs': logits, 'mdl_outs_eval': ev}
    if keep:
        out.update(lang=lang, prop_seg=ps, obj_out=y_obj, mul_in=x_mul, mul_out=y_mul)
    return out


# ---------------------------------------------------------------------------------------------
# SEP concatenation: code/mdl_conc_sep.py:13-217
# ---------------------------------------------------------------------------------------------
def _final_hidden(sd, inp, vocab_size):
    """lang_encode's 'final_hidden' (code/mdl_vog.py:265-279): last layer's final forward|backward hidden state
    (utils/mdl_srl_utils.py:147-162 combine_bidir) through lstm_out_feat_proj -> [B*nv, 256]."""
    words = inp['srl_arg_words_ind']
    B, nv, nsrl, L = words.shape
    flat = words.reshape(B * nv, nsrl * L)
    wm = inp['srl_arg_word_mask'].reshape(B * nv, -1).clone()
    pad = wm == -1
    wm[pad] = 0
    toks = torch.gather(flat, 1, wm)
    toks[pad] = vocab_size
    lens = inp['srl_arg_word_mask_len'].reshape(B * nv)
    toks = toks[:, :int(lens.max())].contiguous()
    emb = F.embedding(toks, sd['lstm_encoder.embed_tokens.weight'])
    packed = torch.nn.utils.rnn.pack_padded_sequence(emb.transpose(0, 1), lens.tolist(), enforce_sorted=False)
    _, (hn, _) = _lstm_from_state_dict(sd)(packed)
    last = torch.cat([hn[-2], hn[-1]], -1)                   # layer 1: forward | backward
    return F.relu(last @ sd['lstm_out_feat_proj.0.weight'].t() + sd['lstm_out_feat_proj.0.bias'])


def vog_forward_sep(sd, inp, nppf, n_heads=3, vocab_size=1000, nfrm0=10):
    """Restated ConcSEP.forward for mdl.name='vog' (code/mdl_conc_sep.py:131-217).

    Every (query b, video c) pair is an independent single-video problem: the object transformer runs over the
    video's own nfrm*nppf proposals (simple_obj_interact with B*ncmp sequences, code/mdl_vog.py:505-516), the
    multimodal transformer over [B*ncmp*nfrm, nsrl*nppf] (conc_encode_item -> conc_encode2, :681-744), i.e. the
    TEMP forward with one video per query applied to the B*ncmp pairs.  On top of that: the video-level verb
    score (:365-398) and the fused per-video score fin_scores (mdl_conc_sep.py:62-117)."""
    B, nv, nsrl, L = inp['srl_arg_words_ind'].shape
    ncmp = inp['new_srl_idxs'].shape[1]
    assert nv == ncmp, 'append_everywhere batches carry the sentence for every video'
    Bq = B * ncmp
    flat = {
        'srl_arg_words_ind': inp['srl_arg_words_ind'].reshape(Bq, 1, nsrl, L),
        'srl_arg_word_mask': inp['srl_arg_word_mask'].reshape(Bq, 1, -1),
        'srl_arg_word_mask_len': inp['srl_arg_word_mask_len'].reshape(Bq, 1),
        'srl_arg_words_capture': inp['srl_arg_words_capture'].reshape(Bq, 1, nsrl, 2),
        'srl_arg_inds_msk': inp['srl_arg_inds_msk'].reshape(Bq, 1, nsrl),
        'pad_region_feature': inp['pad_region_feature'].reshape(Bq, *inp['pad_region_feature'].shape[2:]),
        'seg_feature_for_frms': inp['seg_feature_for_frms'].reshape(Bq, *inp['seg_feature_for_frms'].shape[2:]),
        'pad_proposals': inp['pad_proposals'].reshape(Bq, *inp['pad_proposals'].shape[2:]),
        'new_srl_idxs': inp['new_srl_idxs'].reshape(Bq, 1),
        'num_cmp_msk': inp['num_cmp_msk'].reshape(Bq, 1),
    }
    single = vog_forward(sd, flat, 'temp', nppf, n_heads=n_heads, vocab_size=vocab_size, nfrm0=nfrm0)
    P1 = single['mdl_outs'].shape[-1]
    logits = single['mdl_outs'].view(B, ncmp, nsrl, P1)
    ev = single['mdl_outs_eval'].view(B, ncmp, nsrl, P1)                   # mdl_conc_sep.py:196-208

    # ---- video-level verb score: mean segment feature | sentence encoding -> seg_verb_classf (:365-398)
    seg = F.relu(inp['seg_feature_for_frms'] @ sd['seg_encoder.0.weight'].t() + sd['seg_encoder.0.bias'])
    seg_for_verb = seg.mean(dim=-2)                                        # [B,ncmp,256]
    verb = _final_hidden(sd, inp, vocab_size).view(B, nv, -1)
    sv = torch.cat([verb, seg_for_verb], -1)
    hv = F.relu(sv @ sd['seg_verb_classf.0.weight'].t() + sd['seg_verb_classf.0.bias'])
    vidf = (hv @ sd['seg_verb_classf.2.weight'].t() + sd['seg_verb_classf.2.bias']).squeeze(-1)   # [B,ncmp]

    # ---- fin_scores (mdl_conc_sep.py:62-117, use_vis_msk=True): best box per argument of the UNMASKED
    # sigmoid scores, the verb slot replaced by the video-level score, averaged over the populated slots
    best = torch.sigmoid(logits).max(dim=-1)[0]                            # [B,ncmp,nsrl]
    smsk = inp['srl_arg_inds_msk'].float()
    vm = inp['verb_ind_in_srl'].view(B, ncmp, 1)
    best = best.scatter(2, vm, torch.sigmoid(vidf).unsqueeze(-1))
    best = best * smsk
    cm = inp['num_cmp_msk'].float()
    fin_eval = best.sum(-1) / smsk.sum(-1) * cm
    fin_loss = best * cm.unsqueeze(-1)
    return {'mdl_outs': logits, 'mdl_outs_eval': ev, 'vidf_outs': vidf, 'fin_scores_loss': fin_loss,
            'fin_scores': fin_eval}


def select_boxes_sep(out, pad_proposals, nppf, nfrm0=10):
    """EvaluatorSEP.get_out_results_boxes (code/eval_vsrl_corr.py:162-220): per (argument, video, frame) best
    proposal and its 7-float row; the predicted video is argmax over fin_scores, broadcast over arguments and
    frames.  -> boxes [B,nsrl,ncmp,nfrm,7], scores [B,nsrl,ncmp,nfrm], indexs [B,nsrl,nfrm] int64."""
    ev = out['mdl_outs_eval']
    B, ncmp, nsrl, P1 = ev.shape
    pd = pad_proposals.shape[-1]
    s = ev.transpose(1, 2).contiguous().view(B, nsrl, ncmp, nfrm0, nppf)
    sc, ix = torch.max(s, dim=-1)
    pr = pad_proposals.view(B, 1, ncmp, nfrm0, nppf, pd).expand(B, nsrl, ncmp, nfrm0, nppf, pd)
    bx = torch.gather(pr, -2, ix[..., None, None].expand(B, nsrl, ncmp, nfrm0, 1, pd)).squeeze(-2)
    vid = torch.argmax(out['fin_scores'], dim=-1)
    return {'boxes': bx, 'scores': sc, 'indexs': vid.view(B, 1, 1).expand(B, nsrl, nfrm0).contiguous()}


# ---------------------------------------------------------------------------------------------
# selection: code/eval_vsrl_corr.py:289-345 (TEMP) / :357-424 (SPAT)
# ---------------------------------------------------------------------------------------------
def select_boxes(mdl_outs_eval, pad_proposals, conc_type, ncmp, nppf, nfrm0=10):
    """per-(srl,frame,vid) max / argmax over the nppf proposals, gather their 7-float rows,
    argmax over vids.  Lowest index wins ties (torch.max / argmax semantics).

    -> boxes [B,nsrl,ncmp,nfrm,7], scores [B,nsrl,ncmp,nfrm], indexs [B,nsrl,nfrm] (int64; zeros
    for TEMP, float zeros in the reference :339-341)."""
    B, nv, nsrl, P = mdl_outs_eval.shape
    assert nv == 1
    pd = pad_proposals.shape[-1]
    if conc_type == 'spat':
        s = mdl_outs_eval.view(B, nsrl, nfrm0, ncmp, nppf)
        sc, ix = torch.max(s, dim=-1)                                       # [B,nsrl,nfrm,ncmp]
        pr = pad_proposals.view(B, 1, nfrm0, ncmp, nppf, pd).expand(B, nsrl, nfrm0, ncmp, nppf, pd)
        bx = torch.gather(pr, -2, ix[..., None, None].expand(B, nsrl, nfrm0, ncmp, 1, pd)).squeeze(-2)
        return {'boxes': bx.transpose(2, 3).contiguous(),
                'scores': sc.transpose(2, 3).contiguous(),
                'indexs': sc.argmax(dim=-1)}
    s = mdl_outs_eval.view(B, nsrl, ncmp, nfrm0, nppf)
    sc, ix = torch.max(s, dim=-1)
    pr = pad_proposals.view(B, 1, ncmp, nfrm0, nppf, pd).expand(B, nsrl, ncmp, nfrm0, nppf, pd)
    bx = torch.gather(pr, -2, ix[..., None, None].expand(B, nsrl, ncmp, nfrm0, 1, pd)).squeeze(-2)
    return {'boxes': bx, 'scores': sc, 'indexs': torch.zeros(B, nsrl, nfrm0, dtype=torch.int64)}


# ---------------------------------------------------------------------------------------------
# loss: code/mdl_conc_single.py:180-433 + utils/box_utils.py:54-118
# ---------------------------------------------------------------------------------------------
def bbox_overlaps_batch(anchors, gt_boxes, frm_mask):
    """utils/box_utils.py:61-118 restated: +1 pixel convention, overlaps MULTIPLIED by frm_mask (:108-110),
    zero-area gt boxes -> 0, zero-area anchors -> -1."""
    a, g = anchors[..., :4], gt_boxes[..., :4]
    gx, gy = g[..., 2] - g[..., 0] + 1, g[..., 3] - g[..., 1] + 1
    ax, ay = a[..., 2] - a[..., 0] + 1, a[..., 3] - a[..., 1] + 1
    g_area, a_area = (gx * gy)[:, None, :], (ax * ay)[:, :, None]
    iw = (torch.min(a[:, :, None, 2], g[:, None, :, 2]) - torch.max(a[:, :, None, 0], g[:, None, :, 0]) + 1).clamp(min=0)
    ih = (torch.min(a[:, :, None, 3], g[:, None, :, 3]) - torch.max(a[:, :, None, 1], g[:, None, :, 1]) + 1).clamp(min=0)
    ua = a_area + g_area - iw * ih
    ov = iw * ih / ua
    ov = ov * frm_mask.to(ov.dtype)
    ov = ov.masked_fill(((gx == 1) & (gy == 1))[:, None, :].expand_as(ov), 0)
    ov = ov.masked_fill(((ax == 1) & (ay == 1))[:, :, None].expand_as(ov), -1)
    return ov


def loss_forward(mdl_outs, inp, conc_type, ncmp, nppf, nfrm0=10, loss_lambda=1.0):
    """LossB_SPAT / LossB_TEMP.forward restated (code/mdl_conc_single.py:191-311 TEMP, :342-433 SPAT):
    IoU targets of the target video's proposals against the gt boxes of every SRL argument (> 0.5), BCE with
    logits, mean over the (argument has boxes) x (video valid) mask, times the number of proposals.
    -> {'loss', 'mdl_out_loss'} and the boolean targets [B,1,nsrl,P]."""
    props, gt = inp['pad_proposals'], inp['pad_gt_bboxs']
    frm = inp['pad_frm_mask'] | inp['pad_pnt_mask'].unsqueeze(-1)                   # :229-231
    ov = bbox_overlaps_batch(props[:, :, :5], gt[:, :, :5], frm)
    B, P, K = ov.shape
    idx = torch.arange(P)
    vid = (idx // nppf) % ncmp if conc_type == 'spat' else idx // (P // ncmp)      # :255-263 / :350-363
    ov = ov * (vid[None, :] == inp['target_cmp'].view(B, 1)).to(ov.dtype)[:, :, None]
    sb, sl = inp['srl_boxes'], inp['srl_boxes_lens']                              # [B,1,nsrl,nb]
    nsrl, nb = sb.shape[2], sb.shape[3]
    g = torch.gather(ov.view(B, 1, 1, P, K).expand(B, 1, nsrl, P, K), -1,
                     sb.view(B, 1, nsrl, 1, nb).expand(B, 1, nsrl, P, nb))        # :207-218
    g = g * sl.float().unsqueeze(-2)
    targets = g.max(dim=-1)[0] > 0.5                                              # :226
    tot = torch.nn.functional.binary_cross_entropy_with_logits(mdl_outs, targets.float(), reduction='none')
    bm = inp['srl_arg_boxes_mask'].view(B, 1, nsrl, 1).float() * inp['num_cmp_msk'].view(B, 1, 1, ncmp).float()
    if conc_type == 'spat':                                                      # :399-407
        bm = bm.view(B, 1, nsrl, 1, ncmp, 1).expand(B, 1, nsrl, nfrm0, ncmp, nppf).reshape(B, 1, nsrl, P)
    else:                                                                        # :296-301
        bm = bm.unsqueeze(-1).expand(B, 1, nsrl, ncmp, P // ncmp).reshape(B, 1, nsrl, P)
    sel = torch.masked_select(tot, bm.bool()) if inp['srl_arg_boxes_mask'].max() > 0 else tot
    loss = sel.mean() * tot.size(-1) * loss_lambda
    return {'loss': loss, 'mdl_out_loss': loss, 'targets': targets}


def loss_forward_sep(out, inp, loss_lambda=1.0):
    """LossB_SEP.forward restated (code/mdl_conc_sep.py:236-447): per-video IoU targets (only the target video of
    a query keeps its overlaps, :301-321), BCE with logits masked by num_cmp_msk alone (:341-365; the argument
    mask only decides masked-vs-plain mean), and the verb loss on the video-level logits (:418-434).  'loss' is
    the grounding term only (:436-437).  -> {'loss','mdl_out_loss','verb_loss'} and targets [B,ncmp,nsrl,P1]."""
    props, gt = inp['pad_proposals'], inp['pad_gt_bboxs']
    B, ncmp, P1 = props.shape[:3]
    frm = inp['pad_frm_mask'] | inp['pad_pnt_mask'].unsqueeze(-1)                     # :283-290
    ov = bbox_overlaps_batch(props.reshape(B * ncmp, P1, -1)[:, :, :5], gt.reshape(B * ncmp, *gt.shape[2:])[:, :, :5],
                             frm.reshape(B * ncmp, P1, -1)).view(B, ncmp, P1, -1)
    K = ov.shape[-1]
    onehot = (torch.arange(ncmp).view(1, ncmp) == inp['target_cmp'].view(B, 1)).to(ov.dtype)
    ov = ov * onehot.view(B, ncmp, 1, 1)                                               # :306-314
    sb, sl = inp['srl_boxes'], inp['srl_boxes_lens']
    if sb.shape[1] == 1 and ncmp > 1:
        sb = sb.expand(-1, ncmp, -1, -1)
    nsrl, nb = sb.shape[2], sb.shape[3]
    g = torch.gather(ov.view(B, ncmp, 1, P1, K).expand(B, ncmp, nsrl, P1, K), -1,
                     sb.view(B, ncmp, nsrl, 1, nb).expand(B, ncmp, nsrl, P1, nb))      # :250-262
    g = g * sl.float().unsqueeze(-2)
    targets = g.max(dim=-1)[0] > 0.5
    tot = F.binary_cross_entropy_with_logits(out['mdl_outs'], targets.float(), reduction='none')
    bm = inp['num_cmp_msk'].view(B, ncmp, 1, 1).expand(B, ncmp, nsrl, P1).float()
    tot = tot * bm
    sel = torch.masked_select(tot, bm.bool()) if inp['srl_arg_boxes_mask'].max() > 0 else tot
    mdl_loss = sel.mean() * P1
    vl = F.binary_cross_entropy_with_logits(out['vidf_outs'], inp['verb_cmp'].float(), reduction='none')
    vm = (inp['verb_cross_cmp_msk'].float().sum(-1) > 0)
    verb_loss = torch.masked_select(vl * vm.float(), vm).mean()
    return {'loss': mdl_loss * loss_lambda, 'mdl_out_loss': mdl_loss * loss_lambda, 'verb_loss': verb_loss * loss_lambda,
            'targets': targets}


# ---------------------------------------------------------------------------------------------
# contrastive-sample concatenation: code/dat_loader_simple.py:1067-1207 (SPAT),
# :1231-1292 (TEMP)
# ---------------------------------------------------------------------------------------------
def concat_videos(feat, seg, props, conc_type, nfrm, nppf):
    """Per-video feat [B,ncmp,nfrm*nppf,D], seg [B,ncmp,nfrm,Ds], props [B,ncmp,nfrm*nppf,pdim] -> the concatenated
    single-video tensors.  SPAT: x1,x2 += 720*vid (process_props :1081-1103), rows [vid][frame][prop] ->
    [frame][vid][prop] (reshuffle_boxes :1067-1078), seg [vid][frame] -> [frame][vid] (:1200-1203).
    TEMP: frame id += 10*vid (:1231-1252), order unchanged."""
    B, ncmp, P1, D = feat.shape
    pdim = props.shape[-1]
    v = torch.arange(ncmp, dtype=torch.float32).view(1, ncmp, 1)
    props = props.clone()
    if conc_type == 'spat':
        props[..., 0] = props[..., 0] + v * 720.0
        props[..., 2] = props[..., 2] + v * 720.0
        props = props.view(B, ncmp, nfrm, nppf, pdim).transpose(1, 2).reshape(B, ncmp * P1, pdim)
        feat = feat.view(B, ncmp, nfrm, nppf, D).transpose(1, 2).reshape(B, ncmp * P1, D)
        seg = seg.transpose(1, 2).reshape(B, ncmp * nfrm, seg.shape[-1])
    else:
        props[..., 4] = props[..., 4] + v * 10.0
        props = props.reshape(B, ncmp * P1, pdim)
        feat = feat.reshape(B, ncmp * P1, D)
        seg = seg.reshape(B, ncmp * nfrm, seg.shape[-1])
    return feat.contiguous(), seg.contiguous(), props.contiguous()
