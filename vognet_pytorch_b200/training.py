"""Training step of the VOGNet fusion path: forward that keeps what the backward needs, the analytic backward
through libvog_b200 kernels, and the glue that makes ``model(batch)`` differentiable for the reference trainer
(utils/trn_utils.py:497-505: ``out = mdl(batch); loss = loss_fn(out, batch); loss.backward(); optimizer.step()``).

The reference's backward is torch autograd over its Python forward.  Here the whole model forward is ONE
``torch.autograd.Function`` whose backward runs hand-written kernels in reverse order and returns one gradient per
parameter (``None`` for the three heads the temp/spat forward never reads - what DistributedDataParallel's
``find_unused_parameters=True`` tolerates, code/main_dist.py:75-80):

    logits <- lin2 <- un-regroup <- mult_txf layers <- tokens [vis | lang] <- { obj_txf layers <- prop|seg encoders ,
                                                                             language side: enc <- gather <- proj <- LSTM <- emb }
    relative-position bias: d bias -> (da per box, d b_pe) inside the attention backward -> pe_*_sub_enc weight.
    The boxes carry no gradient (.clone().detach(), code/mdl_vog.py:497,506,624).

Compute modes: 'fp32x' = every product in IEEE fp32 on CUDA cores (gradient parity vs the reference: tests/golden/
grad_cpu_ref.npz); 'tf32' / 'bf16' = tcgen05 GEMMs and attention (training.TcBackend).  Dropout (attention
probabilities, both residual branches, LSTM input / inter-layer / output: code/transformer_code.py:26,31,153;
utils/mdl_srl_utils.py:104,128,150) uses a counter-based Philox stream so the backward regenerates the masks."""
import math

import torch

from . import ops, ops_bwd as ob
from .transformer_code import RelBias


class Tape(dict):
    """Activations the backward needs, by name."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _zeros_like_param(p):
    return torch.zeros_like(p, memory_format=torch.contiguous_format)


class GradSink:
    """name -> gradient tensor (fp32, zero on first touch; kernels accumulate into them).  With `views` (name -> view
    of an optimizer's flat gradient buffer, already zeroed) the kernels write straight into that buffer."""

    def __init__(self, named_params, views=None):
        self.params = dict(named_params)
        self.views = views
        self.g = {}

    def get(self, name):
        t = self.g.get(name)
        if t is None:
            t = self.g[name] = self.views[name] if self.views is not None else _zeros_like_param(self.params[name])
        return t

    def acc(self, name):
        """the tensor a weight-gradient GEMM accumulates into directly (no temporary + add): the flat-buffer view, or
        zeros on first touch"""
        t = self.get(name)
        return t if t.dim() == 2 and t.is_contiguous() else None

    def set(self, name, val):
        if self.views is not None or name in self.g:
            self.get(name).add_(val.view_as(self.params[name]))
        else:
            self.g[name] = val.reshape(self.params[name].shape).contiguous()


# =============================================================================================
# exact-fp32 backend
# =============================================================================================
class F32Backend:
    """Linear algebra of the training step in exact fp32 (vog_sgemm_nt / vog_sgemm_strided / vog_attn_*_f32)."""
    name = 'fp32x'

    def linear(self, x, w, b=None, relu=False, residual=None, key=None, out=None, params=None):
        return ops.sgemm_nt(x, w() if callable(w) else w, b, residual=residual, relu=relu, out=out)

    def lin_dx(self, dy, w, residual=None, key=None, params=None):
        """dy [M,N] @ w [N,K] (+ residual)"""
        w = w() if callable(w) else w
        if residual is None:
            return ob.sgemm(dy, w)
        out = residual.clone() if residual.is_contiguous() else residual.contiguous()
        return ob.sgemm(dy, w, out=out, accumulate=True)

    def lin_dw(self, dy, x, out=None):
        """dy^T @ x -> [N,K]; `out` (fp32 [N,K]): accumulated into, e.g. a view of the flat gradient buffer"""
        if out is None:
            out = torch.zeros(dy.shape[1], x.shape[1], device=dy.device, dtype=torch.float32)
        return ob.sgemm(dy.t(), x, out=out, accumulate=True)


class TcBackend:
    """The same three contractions on tcgen05 with bf16 operands (fp32 accumulate): the skinny language-side GEMMs
    (M = 20 words x B sentences against 8192 x 2048 LSTM matrices) are bound by streaming the weights, which the
    bf16 copies halve and the TMA / tensor-core kernels actually reach.  Weight copies (plain and transposed) are
    cached per parameter version in the model's PackCache under `key`."""
    name = 'bf16'

    def __init__(self, mdl):
        self.packs = mdl._packs()

    @staticmethod
    def lp(x):
        return x if x.dtype == torch.bfloat16 else ops.cast_lp(x, ops.LP_BF16)

    def _w(self, w, key, params, transposed):
        # `w` is a parameter, or a zero-argument callable that derives the matrix from LIVE parameters (the cache
        # re-runs `build` when those change, so it must not close over a stale copy)
        def build():
            m = w() if callable(w) else w
            t = m.detach().t() if transposed else m.detach()
            return t.contiguous().to(torch.bfloat16)

        def into(dst):                       # one fused (transposing) cast-copy into the existing bf16 pack
            m = w() if callable(w) else w
            t = m.detach().t() if transposed else m.detach()
            if dst.shape != t.shape:
                return False
            dst.copy_(t)
        if key is None:
            return build()
        return self.packs.get(('be', key, transposed), params if params is not None else (w,), build, into)

    def linear(self, x, w, b=None, relu=False, residual=None, key=None, out=None, params=None):
        o, _ = ops.tc_gemm(self.lp(x), self._w(w, key, params, False), bias=b, relu=relu, residual=residual, out_f32=out)
        return o

    def lin_dx(self, dy, w, residual=None, key=None, params=None):
        o, _ = ops.tc_gemm(self.lp(dy), self._w(w, key, params, True), residual=residual)
        return o

    def lin_dw(self, dy, x, out=None):
        return ob.tc_gemm_tn(self.lp(dy), self.lp(x), out=out)


# =============================================================================================
# transformer layer: forward keeping activations, backward
# =============================================================================================
def _bias_kw(bias):
    if isinstance(bias, RelBias):
        return dict(bias_mode=ops.BIAS_RANK1, a=bias.a, nbox=bias.nbox, bpe=bias.b)
    if bias is not None:
        return dict(bias_mode=ops.BIAS_DENSE, dense=bias.contiguous())
    return dict(bias_mode=ops.BIAS_NONE)


def _res_ln_f32(branch, x2, ln, dc, p, site):
    """LN(x + drop(branch)) (code/transformer_code.py:30-31); `branch` already holds x + branch when p == 0."""
    pre = branch if p <= 0.0 else ob.dropout(branch, p, dc.seed, site, residual=x2, out=branch)[0]
    return pre, ops.add_layernorm(pre, None, ln.weight, ln.bias, ln.eps)


def stack_forward_f32(ex, x2, Bt, N, bias, be, dc=None, site0=0):
    """EncoderExecutor._run_fp32x with a tape.  x2 [Bt*N, d] -> (y [Bt*N, d], [layer tapes]).  Dropout (dc active):
    the same call sites, seeds and counter-based masks as the tensor-core step (training_tc.stack_forward_tc)."""
    d, H = ex.d, ex.H
    inv_scale = 1.0 / math.sqrt(d)
    tapes = []
    bkw = _bias_kw(bias)
    p = dc.p_tx(ex) if dc is not None else 0.0
    for l, layer in enumerate(ex.stack.layers):
        att, ffn = layer.selfattn, layer.feedforward
        t = Tape(x=x2)
        t.qkv = ops.sgemm_nt(x2, ex._packed(l, layer))
        t.lse = torch.empty(Bt * H * N, device=x2.device, dtype=torch.float32)
        t.seed = dc.site_seed(1000 + site0 + l) if p > 0.0 else 0
        t.o = ops.attn_fwd_f32(t.qkv[:, :d], t.qkv[:, d:2 * d], t.qkv[:, 2 * d:], Bt, N, ex.head_dims, inv_scale,
                               lse=t.lse, drop_p=p, seed=t.seed, **bkw)
        br = ops.sgemm_nt(t.o, att.layer.wo.weight, residual=x2 if p <= 0.0 else None)
        t.pre, t.y = _res_ln_f32(br, x2, att.layernorm, dc, p, site0 + 2 * l)
        t.h = ops.sgemm_nt(t.y, ffn.layer.linear1.weight, ffn.layer.linear1.bias, relu=True)
        br2 = ops.sgemm_nt(t.h, ffn.layer.linear2.weight, ffn.layer.linear2.bias, residual=t.y if p <= 0.0 else None)
        t.pre2, x2 = _res_ln_f32(br2, t.y, ffn.layernorm, dc, p, site0 + 2 * l + 1)
        tapes.append(t)
    return x2, tapes


def stack_backward_f32(ex, prefix, tapes, dout, Bt, N, bias, sink, be, da=None, dbpe=None, dc=None, site0=0):
    """Gradient of a post-LN encoder stack (code/transformer_code.py:84-125,189-241).  dout [Bt*N, d] ->
    d input [Bt*N, d]; parameter gradients into `sink` under `prefix`.encoder.layers.{l}...; da / dbpe accumulate
    the relative-position bias gradients (rank-1 form)."""
    d, H = ex.d, ex.H
    inv_scale = 1.0 / math.sqrt(d)
    bkw = _bias_kw(bias)
    p = dc.p_tx(ex) if dc is not None else 0.0
    for l in reversed(range(len(tapes))):
        layer, t = ex.stack.layers[l], tapes[l]
        att, ffn = layer.selfattn, layer.feedforward
        pl = f'{prefix}.encoder.layers.{l}'
        # ---- feed-forward residual block: out = LN(y + drop(W2 relu(W1 y + b1) + b2))
        dpre2, _ = ob.layernorm_bwd(dout, t.pre2, ffn.layernorm.weight, sink.get(pl + '.feedforward.layernorm.weight'),
                                    sink.get(pl + '.feedforward.layernorm.bias'),
                                    dxsum=sink.get(pl + '.feedforward.layer.linear2.bias') if p <= 0.0 else None,
                                    eps=ffn.layernorm.eps)
        dbr = dpre2
        if p > 0.0:
            dbr = ob.dropout(dpre2, p, dc.seed, site0 + 2 * l + 1)[0]
            ob.colsum_acc(dbr, sink.get(pl + '.feedforward.layer.linear2.bias'))
        sink.set(pl + '.feedforward.layer.linear2.weight', be.lin_dw(dbr, t.h))
        dh = be.lin_dx(dbr, ffn.layer.linear2.weight)
        ob.relu_bwd(dh, t.h, dbias=sink.get(pl + '.feedforward.layer.linear1.bias'), inplace=True)
        sink.set(pl + '.feedforward.layer.linear1.weight', be.lin_dw(dh, t.y))
        dy = be.lin_dx(dh, ffn.layer.linear1.weight, residual=dpre2)
        # ---- attention residual block: y = LN(x + drop(Wo attn(x)))
        dpre, _ = ob.layernorm_bwd(dy, t.pre, att.layernorm.weight, sink.get(pl + '.selfattn.layernorm.weight'),
                                   sink.get(pl + '.selfattn.layernorm.bias'), eps=att.layernorm.eps)
        dbr = dpre if p <= 0.0 else ob.dropout(dpre, p, dc.seed, site0 + 2 * l)[0]
        sink.set(pl + '.selfattn.layer.wo.weight', be.lin_dw(dbr, t.o))
        do = be.lin_dx(dbr, att.layer.wo.weight)
        dqkv, _ = ob.attn_bwd_f32(t.qkv[:, :d], t.qkv[:, d:2 * d], t.qkv[:, 2 * d:], t.o, do, t.lse, Bt, N,
                                  ex.head_dims, inv_scale, da=da, dbpe=dbpe, drop_p=p, seed=t.seed, **bkw)
        dwqkv = be.lin_dw(dqkv, t.x)                                   # [3d, d] = dWq | dWk | dWv
        for i, nm in enumerate(('wq', 'wk', 'wv')):
            sink.set(f'{pl}.selfattn.layer.{nm}.weight', dwqkv[i * d:(i + 1) * d])
        dout = be.lin_dx(dqkv, ex._packed(l, layer), residual=dpre)
    return dout


# =============================================================================================
# whole model
# =============================================================================================
def _check_supported(mdl):
    if mdl.CONC_TYPE == 'sep':
        raise NotImplementedError('vognet_pytorch_b200: training of the SEP concatenation is not built (temp / spat are)')
    if mdl.compute == 'tf32':
        raise NotImplementedError("vognet_pytorch_b200: the training step runs in compute modes 'bf16' (tcgen05) and "
                                  "'fp32x' (exact); 'tf32' is an inference mode")


def dropout_active(mdl):
    """True when the training forward of this configuration draws dropout masks (attention probabilities and the two
    residual branches with mdl.{obj,mul}_tx.attn_drop, code/transformer_code.py:26,31,153; LSTM input / inter-layer /
    output with 0.1, utils/mdl_srl_utils.py:77,104,128,150)."""
    return bool(mdl.train_dropout)


class DropCtx:
    """Dropout bookkeeping of one training forward: ONE host-side seed per call (drawn from torch's CPU generator, so
    torch.manual_seed makes runs repeatable and nothing synchronises with the device) and a fixed id per call site; the
    backward regenerates every mask from (seed, site)."""
    LSTM_IN, LSTM_MID, LSTM_OUT = 1, 2, 3          # utils/mdl_srl_utils.py:128,104,150
    OBJ, MUL = 100, 200                            # + 2*layer (+1 for the feed-forward branch); attention: 1000 + ...

    def __init__(self, mdl, active):
        self.active = bool(active)
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if self.active else 0
        self.p_lstm = 0.1 if self.active else 0.0  # LSTMEncoder defaults dropout_in = dropout_out = 0.1 (:77)

    def site_seed(self, site):
        return (self.seed + site * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF

    def p_tx(self, ex):
        return float(ex.drop) if self.active else 0.0


def _drop(x, dc, p, site, **kw):
    """dropout call site (forward and backward use the same ids)"""
    if p <= 0.0:
        return x
    return ob.dropout(x, p, dc.seed, site, **kw)[0]


def _lstm_params(lstm, l):
    """the eight parameters of layer l: weight_ih, weight_hh, bias_ih, bias_hh, each forward then reverse"""
    names = [f'{n}_l{l}{s}' for n in ('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh') for s in ('', '_reverse')]
    return tuple(getattr(lstm, n) for n in names)


def _lstm_stacked(lstm, l):
    """forward|reverse stacked matrices of layer l (fresh copies) and the parameters they derive from"""
    ps = _lstm_params(lstm, l)
    with torch.no_grad():
        wih = torch.cat([ps[0], ps[1]], 0)
        whh = torch.stack([ps[2], ps[3]], 0).contiguous()
        bias = torch.cat([ps[4] + ps[6], ps[5] + ps[7]], 0)
    return wih, whh, bias, ps


def _lstm_pack(mdl, lstm, l):
    """(wih [8H,in], whh [2,4H,H], bias [8H]) fp32, cached in the model's PackCache and re-derived IN PLACE from the
    live parameters when those change (every training step: one pass over the 150 MB of a layer, no temporaries)
    -> (pack, parameters)"""
    ps = _lstm_params(lstm, l)

    def into(dst):
        wih, whh, bias = dst
        n = ps[0].shape[0]
        wih[:n].copy_(ps[0]); wih[n:].copy_(ps[1])
        whh[0].copy_(ps[2]); whh[1].copy_(ps[3])
        torch.add(ps[4], ps[6], out=bias[:n]); torch.add(ps[5], ps[7], out=bias[n:])
    return mdl._packs().get(('lstm32', l), ps, lambda: _lstm_stacked(lstm, l)[:3], into), ps


def lang_forward(mdl, inp, tp, dc, be=None):
    """Language side of the training forward (code/mdl_vog.py:67-140,250-283; utils/mdl_srl_utils.py:114-169): exact
    fp32 with the F32Backend, bf16 tensor-core GEMMs around the fp32 recurrence with the TcBackend.
    -> lang [B*nsrl, 256] fp32; activations into `tp`."""
    be = be if be is not None else F32Backend()
    tp.lbe = be
    words = inp['srl_arg_words_ind']
    B, nv, nsrl, L = words.shape
    Bq = B * nv
    wm = inp['srl_arg_word_mask'].reshape(Bq, -1).contiguous()
    T = wm.shape[1]
    lens = inp['srl_arg_word_mask_len'].reshape(Bq).contiguous()
    tp.update(Bq=Bq, T=T, lens=lens, wm=wm, words=words.reshape(Bq, nsrl * L).contiguous())
    lstm = mdl.lstm_encoder.lstm
    x = ops.lang_embed(tp.words, wm, mdl.lstm_encoder.embed_tokens.weight, mdl.vocab_size, ops.LP_NONE)
    x = _drop(x, dc, dc.p_lstm, dc.LSTM_IN)
    tp.lstm = []
    for l in range(lstm.num_layers):
        # stacked forward|reverse matrices, re-derived from the live parameters whenever those change; the low-precision
        # copies of the tensor-core backend derive from this pack (a callable: the cache may re-run it later)
        (wih, whh, bias), ps = _lstm_pack(mdl, lstm, l)
        gx = be.linear(x, (lambda l=l: _lstm_pack(mdl, lstm, l)[0][0]) if be.name != 'fp32x' else wih, bias, key=('wih', l),
                       params=ps)
        hout, acts = ops.lstm_layer_fwd(gx, whh, lens, T, Bq, ops.LP_NONE, want_acts=True)
        tp.lstm.append(Tape(x=x, hout=hout, acts=acts, wih=wih, whh=whh, ps=ps))
        last = l == lstm.num_layers - 1
        x = _drop(hout, dc, dc.p_lstm, dc.LSTM_OUT if last else dc.LSTM_MID + 10 * l)
    tp.top = x
    proj, enc = mdl.lstm_out_feat_proj[0], mdl.srl_arg_words_out_enc[0]
    tp.full = be.linear(x, proj.weight, proj.bias, relu=True, key='proj')                         # [T*Bq, 256]
    tp.cap = inp['srl_arg_words_capture'].reshape(Bq, nsrl, 2).contiguous()
    tp.cat = ops.lang_gather(tp.full, tp.cap, T, Bq, ops.LP_NONE)
    tp.enc = be.linear(tp.cat, enc.weight, enc.bias, relu=True, key='enc')
    tp.smsk = inp['srl_arg_inds_msk'].reshape(Bq * nsrl).contiguous()
    lang, _ = ops.mask_rows(tp.enc, tp.smsk)                                                      # [B*nsrl, 256]
    return lang


def lang_backward(mdl, tp, dlang, sink, dc):
    """dlang [B*nsrl, 256] -> parameter gradients of the language side into `sink`."""
    be = tp.lbe
    Bq, T = tp.Bq, tp.T
    denc, _ = ops.mask_rows(dlang, tp.smsk)                                                # mask_rows backward
    ob.relu_bwd(denc, tp.enc, dbias=sink.get('srl_arg_words_out_enc.0.bias'), inplace=True)
    sink.set('srl_arg_words_out_enc.0.weight', be.lin_dw(denc, tp.cat))
    dcat = be.lin_dx(denc, mdl.srl_arg_words_out_enc[0].weight, key='enc')
    dfull = ob.lang_gather_bwd(dcat, tp.cap, T, Bq)
    ob.relu_bwd(dfull, tp.full, dbias=sink.get('lstm_out_feat_proj.0.bias'), inplace=True)
    sink.set('lstm_out_feat_proj.0.weight', be.lin_dw(dfull, tp.top))
    dtop = be.lin_dx(dfull, mdl.lstm_out_feat_proj[0].weight, key='proj')
    _lstm_backward(mdl, tp, dtop, sink, be, dc)


def forward_train_f32(mdl, inp):
    """Training forward in exact fp32.  -> (logits [B,1,nsrl,P], tape)."""
    be = F32Backend()
    dc = DropCtx(mdl, dropout_active(mdl))
    tp = Tape(be=be, dc=dc)
    feat, seg, props = inp['pad_region_feature'], inp['seg_feature_for_frms'], inp['pad_proposals']
    B, P, _ = feat.shape
    ncmp = inp['new_srl_idxs'].shape[1]
    nppf = mdl.num_prop_per_frm
    nvf = seg.shape[1]
    words = inp['srl_arg_words_ind']
    _, nv, nsrl, L = words.shape
    assert nv == 1 and nvf * nppf == P
    tp.update(B=B, P=P, ncmp=ncmp, nppf=nppf, nvf=nvf, nsrl=nsrl)
    lang = lang_forward(mdl, inp, tp, dc)

    # ---- visual side: prop | seg rows (code/mdl_vog.py:291-314; code/mdl_conc_single.py:50-66,156-174)
    pe_ = mdl.prop_encoder[0].out_features
    x0 = torch.empty(B * P, mdl.ps_dim, device=feat.device, dtype=torch.float32)
    tp.feat2, tp.seg2 = feat.reshape(B * P, -1), seg.reshape(B * nvf, -1)
    ops.sgemm_nt(tp.feat2, mdl.prop_encoder[0].weight, mdl.prop_encoder[0].bias, relu=True, out=x0[:, :pe_])
    segf = ops.sgemm_nt(tp.seg2, mdl.seg_encoder[0].weight, mdl.seg_encoder[0].bias, relu=True)
    x0.view(B * nvf, nppf, mdl.ps_dim)[:, :, pe_:] = segf.unsqueeze(1)
    tp.x0 = x0
    props2 = props.reshape(B * P, props.shape[-1])
    tp.props2 = props2

    # ---- object transformer (code/mdl_vog.py:492-523)
    xv = x0
    tp.obj = None
    if mdl.USE_OBJ_TX and mdl.cfg.mdl.obj_tx.to_use:
        otx = mdl.cfg.mdl.obj_tx
        if otx.one_frm:
            nfrm_o, nppf_o = mdl._groups(ncmp)
            Bt_o, N_o, fdiv = B * nfrm_o, nppf_o, float(nfrm_o)
        else:
            Bt_o, N_o, fdiv = B, P, 1.0
        bias = None
        if otx.use_rel:
            a = ops.pe_project(props2, mdl.pe_obj_sub_enc[0].weight, mdl.vid_w, mdl.vid_h, fdiv)
            bias = RelBias(a, mdl.pe_obj_sub_enc[0].bias, N_o)
        xv, tapes = stack_forward_f32(mdl.obj_txf._exec, x0, Bt_o, N_o, bias, be, dc, DropCtx.OBJ)
        tp.obj = Tape(tapes=tapes, Bt=Bt_o, N=N_o, bias=bias, fdiv=fdiv)

    # ---- tokens [vis | lang] regrouped per frame (code/mdl_vog.py:316-344,693-699) + multimodal transformer
    nfrm, nppf2 = mdl._groups(ncmp)
    tp.update(nfrm=nfrm, nppf2=nppf2)
    vis = xv.view(B, nfrm, 1, nppf2, mdl.ps_dim).expand(B, nfrm, nsrl, nppf2, mdl.ps_dim)
    lng = lang.view(B, 1, nsrl, 1, mdl.lang_dim).expand(B, nfrm, nsrl, nppf2, mdl.lang_dim)
    xm = torch.cat([vis, lng], -1).view(B * nfrm * nsrl * nppf2, mdl.vl_dim)
    tp.mul = None
    if mdl.USE_MUL_TX and mdl.cfg.mdl.mul_tx.to_use:
        mtx = mdl.cfg.mdl.mul_tx
        bias = None
        if mtx.use_rel:
            a = ops.pe_project(props2, mdl.pe_mul_sub_enc[0].weight, mdl.vid_w, mdl.vid_h, float(nfrm))
            bias = RelBias(a, mdl.pe_mul_sub_enc[0].bias, nppf2)
        xm, tapes = stack_forward_f32(mdl.mult_txf._exec, xm, B * nfrm, nsrl * nppf2, bias, be, dc, DropCtx.MUL)
        tp.mul = Tape(tapes=tapes, Bt=B * nfrm, N=nsrl * nppf2, bias=bias)
    tp.xm = xm
    # ---- scorer (code/mdl_vog.py:224-230,675-677) + inverse regroup (:724-737)
    tp.h2 = ops.sgemm_nt(xm, mdl.lin2[0].weight, mdl.lin2[0].bias, relu=True)
    lg = ops.sgemm_nt(tp.h2, mdl.lin2[2].weight, mdl.lin2[2].bias)
    logits = lg.view(B, nfrm, nsrl, nppf2).transpose(1, 2).reshape(B, 1, nsrl, P)
    return logits, tp


def _lstm_backward(mdl, tp, dx_top, sink, be, dc):
    """Backward through the stacked bidirectional LSTM.  dx_top [T*Bq, 2H] = gradient of the top layer's output."""
    lstm = mdl.lstm_encoder.lstm
    T, Bq, lens = tp.T, tp.Bq, tp.lens
    if Bq > 8:
        raise NotImplementedError('vognet_pytorch_b200: LSTM backward handles at most 8 sentences per step')
    dout = dx_top
    for l in reversed(range(lstm.num_layers)):
        lt = tp.lstm[l]
        last = l == lstm.num_layers - 1
        dout = _drop(dout, dc, dc.p_lstm, dc.LSTM_OUT if last else dc.LSTM_MID + 10 * l)   # output / inter-layer dropout
        Hh = lstm.hidden_size
        tc = be.name != 'fp32x'
        hprev = ob.lstm_hprev(lt.hout, lens, T, Bq)
        whh_t = mdl._packs().get(('whhT', l), lt.ps, lambda l=l: _lstm_pack(mdl, lstm, l)[0][1].transpose(1, 2).contiguous(),
                                 lambda dst, l=l: dst.copy_(_lstm_pack(mdl, lstm, l)[0][1].transpose(1, 2)))
        dG = ob.lstm_bwd_steps(dout, lt.acts, whh_t, lens, T, Bq, whh=lt.whh)    # [T*Bq, 8H]; activations kept by the forward
        sfx = ('', '_reverse')
        dG_lp = be.lp(dG) if tc else dG
        x_lp = be.lp(lt.x) if tc else lt.x
        db = torch.zeros(8 * Hh, device=dG.device, dtype=torch.float32)
        ob.colsum_acc(dG, db)
        hprev_lp = be.lp(hprev) if tc else hprev
        for d_ in range(2):
            sl = slice(d_ * 4 * Hh, (d_ + 1) * 4 * Hh)
            # the two big weight gradients of a direction accumulate straight into their gradient tensors
            be.lin_dw(dG_lp[:, sl], x_lp, out=sink.get(f'lstm_encoder.lstm.weight_ih_l{l}{sfx[d_]}'))
            sink.set(f'lstm_encoder.lstm.bias_ih_l{l}{sfx[d_]}', db[sl])
            sink.set(f'lstm_encoder.lstm.bias_hh_l{l}{sfx[d_]}', db[sl].clone())
            be.lin_dw(dG_lp[:, sl], hprev_lp[:, d_ * Hh:(d_ + 1) * Hh], out=sink.get(f'lstm_encoder.lstm.weight_hh_l{l}{sfx[d_]}'))
        dout = be.lin_dx(dG_lp, (lambda l=l: _lstm_pack(mdl, lstm, l)[0][0]) if tc else lt.wih, key=('wih', l),
                         params=lt.ps)                                # [T*Bq, in]
    dout = _drop(dout, dc, dc.p_lstm, dc.LSTM_IN)
    ob.lang_embed_bwd(tp.words, tp.wm, dout, mdl.vocab_size, lens, sink.get('lstm_encoder.embed_tokens.weight'))


def backward_train_f32(mdl, tp, dlogits, sink=None, on_lang_done=None):
    """dlogits [B,1,nsrl,P] -> {parameter name: gradient}.  The language side (82 % of the parameters: the LSTM) is
    differentiated right after the multimodal transformer; `on_lang_done()` is called once its gradients are enqueued
    (a data-parallel step starts their all-reduce there, behind the object transformer's backward)."""
    be = tp.be
    sink = sink if sink is not None else GradSink(mdl.named_parameters())
    B, P, nsrl, nfrm, nppf2, nppf, nvf, ncmp = tp.B, tp.P, tp.nsrl, tp.nfrm, tp.nppf2, tp.nppf, tp.nvf, tp.ncmp
    dl = dlogits.reshape(B, nsrl, P).contiguous().float()
    # ---- scorer
    dh2, _ = ob.lin2_bwd(dl, tp.h2, mdl.lin2[2].weight, sink.get('lin2.2.weight').view(-1), sink.get('lin2.2.bias'),
                         sink.get('lin2.0.bias'), nfrm, nsrl, nppf2)
    sink.set('lin2.0.weight', be.lin_dw(dh2, tp.xm))
    dxm = be.lin_dx(dh2, mdl.lin2[0].weight)                           # [M, 768]
    # ---- multimodal transformer
    if tp.mul is not None:
        da = dbpe = None
        if isinstance(tp.mul.bias, RelBias):
            da = torch.zeros_like(tp.mul.bias.a)
            dbpe = sink.get('pe_mul_sub_enc.0.bias')
        dxm = stack_backward_f32(mdl.mult_txf._exec, 'mult_txf', tp.mul.tapes, dxm, tp.mul.Bt, tp.mul.N, tp.mul.bias,
                                 sink, be, da=da, dbpe=dbpe, dc=tp.dc, site0=DropCtx.MUL)
        if da is not None:
            ob.pe_project_bwd(tp.props2, da, sink.get('pe_mul_sub_enc.0.weight'), mdl.vid_w, mdl.vid_h, float(nfrm))
    # ---- tokens -> factors
    dlang = torch.zeros(B * nsrl, mdl.lang_dim, device=dxm.device, dtype=torch.float32)
    dvis = ob.xmul_bwd(dxm.contiguous(), dlang, B, nfrm, nsrl, nppf2, mdl.ps_dim)          # [B*P, 512]
    # ---- language side
    lang_backward(mdl, tp, dlang, sink, tp.dc)
    if on_lang_done is not None:
        on_lang_done()
    # ---- object transformer
    if tp.obj is not None:
        da = dbpe = None
        if isinstance(tp.obj.bias, RelBias):
            da = torch.zeros_like(tp.obj.bias.a)
            dbpe = sink.get('pe_obj_sub_enc.0.bias')
        dvis = stack_backward_f32(mdl.obj_txf._exec, 'obj_txf', tp.obj.tapes, dvis, tp.obj.Bt, tp.obj.N, tp.obj.bias,
                                  sink, be, da=da, dbpe=dbpe, dc=tp.dc, site0=DropCtx.OBJ)
        if da is not None:
            ob.pe_project_bwd(tp.props2, da, sink.get('pe_obj_sub_enc.0.weight'), mdl.vid_w, mdl.vid_h, tp.obj.fdiv)
    # ---- encoders
    pe_ = mdl.prop_encoder[0].out_features
    se_ = mdl.ps_dim - pe_
    dprop, _ = ob.relu_bwd(dvis[:, :pe_], tp.x0[:, :pe_], dbias=sink.get('prop_encoder.0.bias'))
    sink.set('prop_encoder.0.weight', be.lin_dw(dprop, tp.feat2))
    dseg = ob.seg_rep_bwd(dvis, tp.x0, pe_, se_, nppf)                                    # [B*nvf, 256] (ReLU applied)
    ob.colsum_acc(dseg, sink.get('seg_encoder.0.bias'))
    sink.set('seg_encoder.0.weight', be.lin_dw(dseg, tp.seg2))
    return sink.g


class _TrainFn(torch.autograd.Function):
    """The whole model forward as one autograd node: its backward is the hand-written kernel chain above."""

    @staticmethod
    def forward(ctx, mdl, inp, names, *params):
        if mdl.compute == 'fp32x':
            logits, tape = forward_train_f32(mdl, inp)
        else:
            from . import training_tc
            logits, tape = training_tc.forward_train_tc(mdl, inp)
        ctx.mdl, ctx.tape, ctx.names = mdl, tape, names
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        mdl, tape = ctx.mdl, ctx.tape
        if mdl.compute == 'fp32x':
            g = backward_train_f32(mdl, tape, dlogits)
        else:
            from . import training_tc
            g = training_tc.backward_train_tc(mdl, tape, dlogits)
        ctx.tape = None
        return (None, None, None) + tuple(g.get(n) for n in ctx.names)


def forward_train(mdl, inp):
    """model(batch) in .train() mode -> {'mdl_outs': differentiable logits, 'mdl_outs_eval': masked scores}."""
    _check_supported(mdl)
    names, params = zip(*[(n, p) for n, p in mdl.named_parameters() if p.requires_grad])
    logits = _TrainFn.apply(mdl, inp, names, *params)
    B, _, nsrl, P = logits.shape
    with torch.no_grad():
        ev = mdl._mask_outputs(logits.detach(), inp, B, nsrl, inp['new_srl_idxs'].shape[1], mdl.num_prop_per_frm, P)
    return {'mdl_outs': logits, 'mdl_outs_eval': ev['mdl_outs_eval']}
