"""Training step of the VOGNet fusion path on the tcgen05 tensor cores (compute mode 'bf16').

Same structure as ``training`` (one autograd node, analytic backward), with every large contraction on the tensor
cores:

    forward   QKV / Wo / FFN / encoder / scorer GEMMs   vog_tc_gemm*            (bf16 operands, fp32 accumulate)
              attention                                  vog_tc_attn_fwd_train   (keeps the row log-sum-exp; dropout on P)
    backward  input gradients  dX = dY . W               vog_tc_gemm on a transposed bf16 copy of the weight
              weight gradients dW = dY^T . X             vog_tc_gemm_tn          (fp32 accumulation into the gradient)
              attention                                  vog_tc_attn_bwd         (P / dS recomputed from the lse)
              LayerNorm, ReLU, dropout, glue             element-wise kernels shared with the exact-fp32 backend

The residual stream, LayerNorm statistics, softmax statistics, bias / LayerNorm gradients and every accumulation stay
fp32; the language side (2-layer bi-LSTM over <= 20 words, a few GFLOP) runs the exact-fp32 kernels in both modes.
Reference: utils/trn_utils.py:497-505 (mdl(batch) -> loss.backward()), code/transformer_code.py:21-31,136-241,
code/mdl_vog.py:291-344,492-523,595-744."""
import math

import torch

from . import ops, ops_bwd as ob
from .training import DropCtx, GradSink, Tape, TcBackend, dropout_active, lang_backward, lang_forward
from .transformer_code import RelBias

KIND = ops.LP_BF16


def _bf16_t(w):
    """transposed bf16 copy of a (packed) weight: the B operand of the input-gradient GEMM dX = dY . W"""
    return w.detach().t().contiguous().to(torch.bfloat16)


def _packed_t(ex, l, layer):
    att, ffn = layer.selfattn.layer, layer.feedforward.layer
    ws = (att.wq.weight, att.wk.weight, att.wv.weight, att.wo.weight, ffn.linear1.weight, ffn.linear2.weight)

    def build():
        w = ex._packed_tc(l, layer, KIND)
        return dict(wqkv_t=_bf16_t(w['wqkv']), wo_t=_bf16_t(w['wo']), w1_t=_bf16_t(w['w1']), w2_t=_bf16_t(w['w2']))
    return ex._cached(('tcT', l), ws, build)


def _bias_kw(bias):
    if isinstance(bias, RelBias):
        return dict(bias_mode=ops.BIAS_RANK1, a=bias.a, nbox=bias.nbox, bpe=bias.b)
    if bias is not None:
        raise NotImplementedError("vognet_pytorch_b200: a dense x_pe trains in the 'fp32x' mode only")
    return dict(bias_mode=ops.BIAS_NONE)


def _residual_ln(branch, x2, ln, dc, p, site):
    """LN(x + drop(branch)) of a ResidualBlock (code/transformer_code.py:30-31).  `branch` already holds x + branch
    when p == 0 (residual folded into the GEMM epilogue).  -> (pre, y fp32, y bf16)"""
    pre = branch if p <= 0.0 else ob.dropout(branch, p, dc.seed, site, residual=x2, out=branch)[0]
    M, d = pre.shape
    y = torch.empty(M, d, device=pre.device, dtype=torch.float32)
    y_lp = torch.empty(M, d, device=pre.device, dtype=torch.bfloat16)
    ops.add_layernorm(pre, None, ln.weight, ln.bias, ln.eps, out=y, out_lp=y_lp, lp_kind=KIND)
    return pre, y, y_lp


def stack_forward_tc(ex, x2, x_lp, Bt, N, bias, dc, site0):
    """Post-LN encoder stack, training forward.  x2 [Bt*N, d] fp32 + bf16 copy -> (y, y_lp, [layer tapes])."""
    d, H, dhp = ex.d, ex.H, ex.dhp
    inv_scale = 1.0 / math.sqrt(d)
    bkw = _bias_kw(bias)
    p = dc.p_tx(ex)
    tapes = []
    for l, layer in enumerate(ex.stack.layers):
        att, ffn = layer.selfattn, layer.feedforward
        w = ex._packed_tc(l, layer, KIND)
        t = Tape(x_lp=x_lp)
        t.q, t.k, t.v = ops.tc_gemm_qkv(x_lp, w['wqkv'], Bt, N, H, dhp)
        t.o_lp, t.lse = ob.tc_attn_fwd_train(t.q, t.k, t.v, N, ex.head_dims, inv_scale, drop_p=p,
                                             seed=dc.site_seed(1000 + site0 + l), **bkw)
        br, _ = ops.tc_gemm(t.o_lp, w['wo'], residual=x2 if p <= 0.0 else None)
        t.pre, y, t.y_lp = _residual_ln(br, x2, att.layernorm, dc, p, site0 + 2 * l)
        _, t.h_lp = ops.tc_gemm(t.y_lp, w['w1'], bias=ffn.layer.linear1.bias, relu=True, lp_kind=KIND, want_f32=False)
        br2, _ = ops.tc_gemm(t.h_lp, w['w2'], bias=ffn.layer.linear2.bias, residual=y if p <= 0.0 else None)
        t.pre2, x2, x_lp = _residual_ln(br2, y, ffn.layernorm, dc, p, site0 + 2 * l + 1)
        tapes.append(t)
    return x2, x_lp, tapes


def _unpad_rows(wp, ex, which):
    """rows of a [3*H*dhp, d] padded-head gradient that belong to projection `which` -> [d, d]"""
    H, dhp = ex.H, ex.dhp
    return torch.cat([wp[(which * H + h) * dhp:(which * H + h) * dhp + dh] for h, dh in enumerate(ex.head_dims)], 0)


def _unpad_cols(wp, ex):
    dhp = ex.dhp
    return torch.cat([wp[:, h * dhp:h * dhp + dh] for h, dh in enumerate(ex.head_dims)], 1)


def stack_backward_tc(ex, prefix, tapes, dout, Bt, N, bias, sink, dc, site0, da=None, dbpe=None):
    """Gradient of the stack: dout [Bt*N, d] fp32 -> d input [Bt*N, d] fp32; parameter gradients into `sink`."""
    d = ex.d
    inv_scale = 1.0 / math.sqrt(d)
    bkw = _bias_kw(bias)
    p = dc.p_tx(ex)
    for l in reversed(range(len(tapes))):
        layer, t = ex.stack.layers[l], tapes[l]
        att, ffn = layer.selfattn, layer.feedforward
        wt = _packed_t(ex, l, layer)
        pl = f'{prefix}.encoder.layers.{l}'
        # ---- feed-forward block: out = LN(y + drop(W2 relu(W1 y + b1) + b2))
        dpre2, dbr_lp = ob.layernorm_bwd(dout, t.pre2, ffn.layernorm.weight, sink.get(pl + '.feedforward.layernorm.weight'),
                                         sink.get(pl + '.feedforward.layernorm.bias'),
                                         dxsum=sink.get(pl + '.feedforward.layer.linear2.bias') if p <= 0.0 else None,
                                         eps=ffn.layernorm.eps, lp_kind=KIND)
        if p > 0.0:          # gradient of the branch = dropped gradient of the sum; its column sums = d b2
            dbr, dbr_lp = ob.dropout(dpre2, p, dc.seed, site0 + 2 * l + 1, lp_kind=KIND)
            ob.colsum_acc(dbr, sink.get(pl + '.feedforward.layer.linear2.bias'))
        sink.set(pl + '.feedforward.layer.linear2.weight', ob.tc_gemm_tn(dbr_lp, t.h_lp))
        dh, _ = ops.tc_gemm(dbr_lp, wt['w2_t'])
        _, dh_lp = ob.relu_bwd(dh, t.h_lp, dbias=sink.get(pl + '.feedforward.layer.linear1.bias'), lp_kind=KIND,
                               want_f32=False)
        sink.set(pl + '.feedforward.layer.linear1.weight', ob.tc_gemm_tn(dh_lp, t.y_lp))
        dy, _ = ops.tc_gemm(dh_lp, wt['w1_t'], residual=dpre2)
        # ---- attention block: y = LN(x + drop(Wo attn(x)))
        dpre, dbr_lp = ob.layernorm_bwd(dy, t.pre, att.layernorm.weight, sink.get(pl + '.selfattn.layernorm.weight'),
                                        sink.get(pl + '.selfattn.layernorm.bias'), eps=att.layernorm.eps, lp_kind=KIND)
        if p > 0.0:
            _, dbr_lp = ob.dropout(dpre, p, dc.seed, site0 + 2 * l, lp_kind=KIND, want_f32=False)
        sink.set(pl + '.selfattn.layer.wo.weight', _unpad_cols(ob.tc_gemm_tn(dbr_lp, t.o_lp), ex))
        _, do_lp = ops.tc_gemm(dbr_lp, wt['wo_t'], lp_kind=KIND, want_f32=False)
        dqkv = ob.tc_attn_bwd(t.q, t.k, t.v, t.o_lp, do_lp, t.lse, N, ex.head_dims, inv_scale, da=da, dbpe=dbpe,
                              drop_p=p, seed=dc.site_seed(1000 + site0 + l), **bkw)
        dwqkv = ob.tc_gemm_tn(dqkv, t.x_lp)                               # [3*H*dhp, d], padded head slots
        for i, nm in enumerate(('wq', 'wk', 'wv')):
            sink.set(f'{pl}.selfattn.layer.{nm}.weight', _unpad_rows(dwqkv, ex, i))
        dout, _ = ops.tc_gemm(dqkv, wt['wqkv_t'], residual=dpre)
    return dout


def forward_train_tc(mdl, inp):
    """Training forward, tensor-core path.  -> (logits [B,1,nsrl,P], tape)."""
    if mdl.compute != 'bf16':
        raise NotImplementedError("vognet_pytorch_b200: the tensor-core training step is compute mode 'bf16'")
    dc = DropCtx(mdl, dropout_active(mdl))
    tp = Tape(dc=dc)
    feat, seg, props = inp['pad_region_feature'], inp['seg_feature_for_frms'], inp['pad_proposals']
    B, P, _ = feat.shape
    ncmp = inp['new_srl_idxs'].shape[1]
    nppf = mdl.num_prop_per_frm
    nvf = seg.shape[1]
    _, nv, nsrl, _ = inp['srl_arg_words_ind'].shape
    assert nv == 1 and nvf * nppf == P
    tp.update(B=B, P=P, ncmp=ncmp, nppf=nppf, nvf=nvf, nsrl=nsrl)
    dev = feat.device
    lang = lang_forward(mdl, inp, tp, dc, TcBackend(mdl))

    # ---- visual side: prop | seg rows written by the two encoder GEMMs (the seg half replicated over the proposals)
    pe_ = mdl.prop_encoder[0].out_features
    x0 = torch.empty(B * P, mdl.ps_dim, device=dev, dtype=torch.float32)
    x0_lp = torch.empty(B * P, mdl.ps_dim, device=dev, dtype=torch.bfloat16)
    tp.feat_lp = ops.cast_lp(feat.reshape(B * P, -1), KIND)
    tp.seg2 = seg.reshape(B * nvf, -1)
    ops.tc_gemm(tp.feat_lp, mdl._lp_weight('prop', mdl.prop_encoder[0].weight, KIND), bias=mdl.prop_encoder[0].bias,
                relu=True, out_f32=x0[:, :pe_], out_lp=x0_lp[:, :pe_])
    tp.seg_lp = ops.cast_lp(tp.seg2, KIND)
    ops.tc_gemm(tp.seg_lp, mdl._lp_weight('seg', mdl.seg_encoder[0].weight, KIND),
                bias=mdl.seg_encoder[0].bias, relu=True, out_f32=x0[:, pe_:], out_lp=x0_lp[:, pe_:], rep=nppf)
    tp.x0 = x0
    props2 = props.reshape(B * P, props.shape[-1])
    tp.props2 = props2

    xv, xv_lp = x0, x0_lp
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
        xv, xv_lp, tapes = stack_forward_tc(mdl.obj_txf._exec, x0, x0_lp, Bt_o, N_o, bias, dc, DropCtx.OBJ)
        tp.obj = Tape(tapes=tapes, Bt=Bt_o, N=N_o, bias=bias, fdiv=fdiv)

    # ---- tokens [vis | lang] (materialised for the training step: the weight gradient of the first projection
    #      contracts over them) + multimodal transformer
    nfrm, nppf2 = mdl._groups(ncmp)
    tp.update(nfrm=nfrm, nppf2=nppf2)
    xm, xm_lp = ops.build_xmul(xv.contiguous(), lang, B, nfrm, nsrl, nppf2, KIND)
    tp.mul = None
    if mdl.USE_MUL_TX and mdl.cfg.mdl.mul_tx.to_use:
        mtx = mdl.cfg.mdl.mul_tx
        bias = None
        if mtx.use_rel:
            a = ops.pe_project(props2, mdl.pe_mul_sub_enc[0].weight, mdl.vid_w, mdl.vid_h, float(nfrm))
            bias = RelBias(a, mdl.pe_mul_sub_enc[0].bias, nppf2)
        xm, xm_lp, tapes = stack_forward_tc(mdl.mult_txf._exec, xm, xm_lp, B * nfrm, nsrl * nppf2, bias, dc, DropCtx.MUL)
        tp.mul = Tape(tapes=tapes, Bt=B * nfrm, N=nsrl * nppf2, bias=bias)
    tp.xm_lp = xm_lp
    # ---- scorer + inverse regroup
    h2, tp.h2_lp = ops.tc_gemm(xm_lp, mdl._lp_weight('lin2', mdl.lin2[0].weight, KIND), bias=mdl.lin2[0].bias, relu=True,
                               lp_kind=KIND)
    logits, _ = ops.lin2_tail(h2, mdl.lin2[2].weight, mdl.lin2[2].bias, inp['srl_arg_inds_msk'].reshape(B, nsrl),
                              inp['num_cmp_msk'], B, nfrm, nsrl, nppf2, ncmp, nppf, mdl.num_sampled_frm,
                              mdl.CONC_TYPE == 'spat')
    return logits, tp


def backward_train_tc(mdl, tp, dlogits, sink=None, on_lang_done=None):
    """dlogits [B,1,nsrl,P] -> {parameter name: gradient (fp32)}; `sink` / `on_lang_done` as in training.backward_train_f32."""
    dc = tp.dc
    sink = sink if sink is not None else GradSink(mdl.named_parameters())
    B, P, nsrl, nfrm, nppf2, nppf = tp.B, tp.P, tp.nsrl, tp.nfrm, tp.nppf2, tp.nppf
    dl = dlogits.reshape(B, nsrl, P).contiguous().float()
    # ---- scorer
    _, dh2_lp = ob.lin2_bwd(dl, tp.h2_lp, mdl.lin2[2].weight, sink.get('lin2.2.weight').view(-1), sink.get('lin2.2.bias'),
                            sink.get('lin2.0.bias'), nfrm, nsrl, nppf2, want_f32=False, lp_kind=KIND)
    sink.set('lin2.0.weight', ob.tc_gemm_tn(dh2_lp, tp.xm_lp))
    w1t = mdl._packs().get(('lin2T', KIND), (mdl.lin2[0].weight,), lambda: _bf16_t(mdl.lin2[0].weight))
    dxm, _ = ops.tc_gemm(dh2_lp, w1t)                                           # [M, 768] fp32
    # ---- multimodal transformer
    if tp.mul is not None:
        da = dbpe = None
        if isinstance(tp.mul.bias, RelBias):
            da = torch.zeros_like(tp.mul.bias.a)
            dbpe = sink.get('pe_mul_sub_enc.0.bias')
        dxm = stack_backward_tc(mdl.mult_txf._exec, 'mult_txf', tp.mul.tapes, dxm, tp.mul.Bt, tp.mul.N, tp.mul.bias, sink,
                                dc, DropCtx.MUL, da=da, dbpe=dbpe)
        if da is not None:
            ob.pe_project_bwd(tp.props2, da, sink.get('pe_mul_sub_enc.0.weight'), mdl.vid_w, mdl.vid_h, float(nfrm))
    # ---- tokens -> factors
    dlang = torch.zeros(B * nsrl, mdl.lang_dim, device=dxm.device, dtype=torch.float32)
    dvis = ob.xmul_bwd(dxm.contiguous(), dlang, B, nfrm, nsrl, nppf2, mdl.ps_dim)          # [B*P, 512]
    # ---- language side (exact fp32), before the object transformer: see training.backward_train_f32
    lang_backward(mdl, tp, dlang, sink, dc)
    if on_lang_done is not None:
        on_lang_done()
    # ---- object transformer
    if tp.obj is not None:
        da = dbpe = None
        if isinstance(tp.obj.bias, RelBias):
            da = torch.zeros_like(tp.obj.bias.a)
            dbpe = sink.get('pe_obj_sub_enc.0.bias')
        dvis = stack_backward_tc(mdl.obj_txf._exec, 'obj_txf', tp.obj.tapes, dvis, tp.obj.Bt, tp.obj.N, tp.obj.bias, sink,
                                 dc, DropCtx.OBJ, da=da, dbpe=dbpe)
        if da is not None:
            ob.pe_project_bwd(tp.props2, da, sink.get('pe_obj_sub_enc.0.weight'), mdl.vid_w, mdl.vid_h, tp.obj.fdiv)
    # ---- encoders
    pe_ = mdl.prop_encoder[0].out_features
    se_ = mdl.ps_dim - pe_
    _, dprop_lp = ob.relu_bwd(dvis[:, :pe_], tp.x0[:, :pe_], dbias=sink.get('prop_encoder.0.bias'), lp_kind=KIND,
                              want_f32=False)
    sink.set('prop_encoder.0.weight', ob.tc_gemm_tn(dprop_lp, tp.feat_lp))
    dseg = ob.seg_rep_bwd(dvis, tp.x0, pe_, se_, nppf)                                    # [B*nvf, 256] (ReLU applied)
    ob.colsum_acc(dseg, sink.get('seg_encoder.0.bias'))
    sink.set('seg_encoder.0.weight', ob.tc_gemm_tn(ops.cast_lp(dseg, KIND), tp.seg_lp))
    return sink.g
