"""Operator-level boundary: ``Transformer`` / ``RelTransformer`` with the reference's constructor
signature, ``forward`` signature and state_dict key names (code/transformer_code.py:244-279), so
reference checkpoints load with ``strict=True``:

    encoder.layers.{l}.selfattn.layer.{wq,wk,wv,wo}.weight      [d,d]   (no bias, :169-172)
    encoder.layers.{l}.selfattn.layernorm.{weight,bias}         [d]
    encoder.layers.{l}.feedforward.layer.linear1.{weight,bias}  [d/2,d]
    encoder.layers.{l}.feedforward.layer.linear2.{weight,bias}  [d,d/2]
    encoder.layers.{l}.feedforward.layernorm.{weight,bias}      [d]

The modules below are parameter containers only; the arithmetic of a layer
(post-LN residual blocks :21-31, multi-head attention with the bias added before the
1/sqrt(d_model) scaling :136-160,176-186, FFN :73-81) runs in the CUDA kernels of libvog_b200
driven by ``EncoderExecutor``.  ``x_pe`` may be the reference's dense [Bt,N,N,H] tensor or a
``RelBias`` descriptor (rank-1 factorisation; never materialises N x N).
"""
import math

import torch
from torch import nn

from . import ops
from .packing import PackCache

COMPUTE_MODES = ('fp32x', 'tf32', 'bf16')


class RelBias:
    """bias[bt,i,j,h] = relu(a[bt*nbox + i % nbox, h] - a[bt*nbox + j % nbox, h] + b[h])

    ``a`` = pe_enc.weight . normalised boxes (ops.pe_project), ``b`` = pe_enc.bias.  Equal to the
    reference's Linear(5,H)+ReLU on pairwise box differences tiled nsrl x nsrl
    (code/mdl_vog.py:477-488) because the Linear is applied to p_i - p_j."""

    def __init__(self, a, b, nbox, ak=None, ak_sig=None):
        self.a, self.b, self.nbox = a, b, int(nbox)
        # optional: the per-key factors already expanded for ONE attention geometry by ops.pe_project_expand
        # (ak_sig = (Bt, N, H, d_model)); the tensor-core executor then skips the attention's expansion pre-kernel
        self.ak, self.ak_sig = ak, ak_sig


class FactoredTokens:
    """The multimodal transformer's input in factored form: token (bt, s, p) of per-frame sequence
    bt = b*nfrm + f is [vis[bt*nppf2 + p] | lang[b*nsrl + s]] (code/mdl_vog.py:316-344,693-699).
    vis [Bt*nppf2, dv] fp32 + low-precision copy, lang [B*nsrl, dl] fp32 + low-precision copy.  The
    executor projects the two factors separately and reads the residual from them, so the
    [Bt, nsrl*nppf2, dv+dl] matrix is never written."""

    def __init__(self, vis, vis_lp, lang, lang_lp, nfrm, nsrl, nppf2):
        self.vis, self.vis_lp, self.lang, self.lang_lp = vis, vis_lp, lang, lang_lp
        self.nfrm, self.nsrl, self.nppf2 = int(nfrm), int(nsrl), int(nppf2)
        self.Bt = vis.shape[0] // self.nppf2
        self.N = self.nsrl * self.nppf2
        self.d = vis.shape[1] + lang.shape[1]


class _Heads(nn.Module):
    def __init__(self, d, n_heads):
        super().__init__()
        self.n_heads = n_heads
        self.wq = nn.Linear(d, d, bias=False)
        self.wk = nn.Linear(d, d, bias=False)
        self.wv = nn.Linear(d, d, bias=False)
        self.wo = nn.Linear(d, d, bias=False)


class _FFN(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.linear1 = nn.Linear(d, f)
        self.linear2 = nn.Linear(f, d)


class _PostLN(nn.Module):
    def __init__(self, layer, d, drop):
        super().__init__()
        self.layer = layer
        self.dropout = nn.Dropout(drop)
        self.layernorm = nn.LayerNorm(d)


class _Layer(nn.Module):
    def __init__(self, d, f, n_heads, drop):
        super().__init__()
        self.selfattn = _PostLN(_Heads(d, n_heads), d, drop)
        self.feedforward = _PostLN(_FFN(d, f), d, drop)


class _Stack(nn.Module):
    def __init__(self, d, f, n_layers, n_heads, drop):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(d, f, n_heads, drop) for _ in range(n_layers)])
        self.dropout = nn.Dropout(drop)


class EncoderExecutor:
    """Runs a ``_Stack`` on [Bt,N,d] CUDA input through libvog_b200."""

    def __init__(self, stack, d_model, n_heads, drop_ratio):
        self.stack = stack
        self.d = d_model
        self.H = n_heads
        self.drop = drop_ratio
        self.head_dims = ops.chunk_sizes(d_model, n_heads)
        self.dhp = ops.round_up(max(self.head_dims), 64)
        self._pack = PackCache()

    # -- packed weights, rebuilt whenever a parameter was updated in place or replaced ---------
    def _cached(self, key, params, build):
        return self._pack.get(key, params, build)

    def _packed(self, l, layer):
        att = layer.selfattn.layer
        ws = (att.wq.weight, att.wk.weight, att.wv.weight)
        return self._cached(('qkv32', l), ws, lambda: torch.cat(ws, 0).detach().contiguous())

    def _packed_tc(self, l, layer, kind):
        """Tensor-core operands of one layer: per-head zero-padded Wq|Wk|Wv rows ([3*H*dhp, d]), Wo
        with matching zero-padded columns ([d, H*dhp]), FFN weights - all in bf16 or tf32-rounded."""
        att, ffn = layer.selfattn.layer, layer.feedforward.layer
        ws = (att.wq.weight, att.wk.weight, att.wv.weight, att.wo.weight, ffn.linear1.weight,
              ffn.linear2.weight)
        dhp = self.dhp

        def lp(t):
            return ops.cast_lp(t.detach().float().contiguous(), kind)

        def build():
            wqkv, wo = ops.pack_weights(ws[0], ws[1], ws[2], ws[3], self.head_dims, dhp, kind)
            return dict(wqkv=wqkv, wo=wo, w1=lp(ws[4]), w2=lp(ws[5]))
        return self._cached(('tc', l, kind), ws, build)

    def _bias_args(self, bias, Bt, N):
        mode, a, bpe, nbox, dense = ops.BIAS_NONE, None, None, 0, None
        if isinstance(bias, RelBias):
            mode, a, bpe, nbox = ops.BIAS_RANK1, bias.a, bias.b, bias.nbox
        elif bias is not None:
            if tuple(bias.shape) != (Bt, N, N, self.H):
                raise ValueError(f'x_pe must be [{Bt},{N},{N},{self.H}], got {tuple(bias.shape)}')
            mode, dense = ops.BIAS_DENSE, bias.contiguous()
        return dict(bias_mode=mode, a=a, nbox=nbox, bpe=bpe, dense=dense)

    def run(self, x, bias, compute, training=False, x_lp=None, want_lp=False):
        """x [Bt,N,d] fp32 (+ optional low-precision copy x_lp) -> y [Bt,N,d] fp32, or (y, y_lp)."""
        if training and self.drop > 0:
            raise NotImplementedError('vognet_pytorch_b200: forward-only build (dropout/backward are '
                                      'scheduled next, SURVEY.md section 8f); call .eval()')
        if compute not in COMPUTE_MODES:
            raise ValueError(f'compute must be one of {COMPUTE_MODES}')
        Bt, N, d = x.shape
        if d != self.d:
            raise ValueError(f'expected last dim {self.d}, got {d}')
        if x.dtype != torch.float32:
            raise TypeError('activations must be float32')
        x2 = x.reshape(Bt * N, d).contiguous()
        bkw = self._bias_args(bias, Bt, N)
        inv_scale = 1.0 / math.sqrt(d)          # sqrt(d_model), not sqrt(d_head): :132,:195
        if compute == 'fp32x':
            y = self._run_fp32x(x2, Bt, N, bkw, inv_scale)
            return (y.view(Bt, N, d), None) if want_lp else y.view(Bt, N, d)
        kind = ops.LP_BF16 if compute == 'bf16' else ops.LP_TF32
        if x_lp is None:
            x_lp = ops.cast_lp(x2, kind)
        self._use_expanded(bkw, bias, Bt, N)
        y, y_lp = self._run_tc(x2, x_lp.reshape(Bt * N, d), Bt, N, bkw, inv_scale, kind)
        return (y.view(Bt, N, d), y_lp) if want_lp else y.view(Bt, N, d)

    def _use_expanded(self, bkw, bias, Bt, N):
        """tensor-core modes: hand the attention the key factors the caller expanded for exactly this geometry"""
        if isinstance(bias, RelBias) and bias.ak is not None and bias.ak_sig == (Bt, N, self.H, self.d):
            bkw['ak'] = bias.ak

    def _run_fp32x(self, x2, Bt, N, bkw, inv_scale):
        d = self.d
        for l, layer in enumerate(self.stack.layers):
            att, ffn = layer.selfattn, layer.feedforward
            qkv = ops.sgemm_nt(x2, self._packed(l, layer))
            o = ops.attn_fwd_f32(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], Bt, N, self.head_dims,
                                 inv_scale, **bkw)
            pre = ops.sgemm_nt(o, att.layer.wo.weight, residual=x2)
            y = ops.add_layernorm(pre, None, att.layernorm.weight, att.layernorm.bias, att.layernorm.eps)
            h = ops.sgemm_nt(y, ffn.layer.linear1.weight, ffn.layer.linear1.bias, relu=True)
            pre2 = ops.sgemm_nt(h, ffn.layer.linear2.weight, ffn.layer.linear2.bias, residual=y)
            x2 = ops.add_layernorm(pre2, None, ffn.layernorm.weight, ffn.layernorm.bias, ffn.layernorm.eps)
        return x2

    def run_factored(self, ft, bias, compute, need_f32=True):
        """tensor-core path on a FactoredTokens input -> (y [Bt,N,d] fp32 or None, y_lp)."""
        if compute not in ('tf32', 'bf16'):
            raise ValueError('run_factored: tensor-core compute modes only')
        if ft.d != self.d:
            raise ValueError(f'expected token width {self.d}, got {ft.d}')
        bkw = self._bias_args(bias, ft.Bt, ft.N)
        self._use_expanded(bkw, bias, ft.Bt, ft.N)
        kind = ops.LP_BF16 if compute == 'bf16' else ops.LP_TF32
        y, y_lp = self._run_tc(None, None, ft.Bt, ft.N, bkw, 1.0 / math.sqrt(self.d), kind, ft=ft,
                               need_f32=need_f32)
        return (y.view(ft.Bt, ft.N, self.d) if y is not None else None), y_lp

    def _run_tc(self, x2, x_lp, Bt, N, bkw, inv_scale, kind, ft=None, need_f32=True):
        """tcgen05 path.  GEMM operands in `kind` (bf16 / tf32-rounded), attention operands always
        bf16, residual stream / LayerNorm / softmax statistics in fp32."""
        M, d, H, dhp = Bt * N, self.d, self.H, self.dhp
        lp_dtype = torch.bfloat16 if kind == ops.LP_BF16 else torch.float32
        for l, layer in enumerate(self.stack.layers):
            att, ffn = layer.selfattn, layer.feedforward
            w = self._packed_tc(l, layer, kind)
            fact = ft is not None and l == 0
            if fact:
                dv = ft.vis.shape[1]
                lq, _ = ops.tc_gemm(ft.lang_lp, w['wqkv'][:, dv:])          # language rows projected once
                q, k, v = ops.tc_gemm_qkv_factored(ft.vis_lp, w['wqkv'][:, :dv], lq, Bt, ft.nfrm, ft.nsrl,
                                                    ft.nppf2, H, dhp)
            else:
                q, k, v = ops.tc_gemm_qkv(x_lp, w['wqkv'], Bt, N, H, dhp)
            o_lp = ops.tc_attn_fwd(q, k, v, N, self.head_dims, inv_scale, out_kind=kind, **bkw)
            if fact:
                pre, _ = ops.tc_gemm_gres(o_lp, w['wo'], ft.vis, ft.lang, ft.nfrm, ft.nsrl, ft.nppf2)
            else:
                pre, _ = ops.tc_gemm(o_lp, w['wo'], residual=x2)
            y = torch.empty(M, d, device=pre.device, dtype=torch.float32)
            y_lp = torch.empty(M, d, device=pre.device, dtype=lp_dtype)
            ops.add_layernorm(pre, None, att.layernorm.weight, att.layernorm.bias, att.layernorm.eps,
                              out=y, out_lp=y_lp, lp_kind=kind)
            _, h_lp = ops.tc_gemm(y_lp, w['w1'], bias=ffn.layer.linear1.bias, relu=True, lp_kind=kind,
                                  want_f32=False)
            pre2, _ = ops.tc_gemm(h_lp, w['w2'], bias=ffn.layer.linear2.bias, residual=y)
            last = l == len(self.stack.layers) - 1
            # the fp32 copy of the stack's output is skipped when the caller only feeds a GEMM with it
            x2 = torch.empty(M, d, device=pre.device, dtype=torch.float32) if (need_f32 or not last) else None
            x_lp = torch.empty(M, d, device=pre.device, dtype=lp_dtype)
            ops.add_layernorm(pre2, None, ffn.layernorm.weight, ffn.layernorm.bias, ffn.layernorm.eps,
                              out=x2, out_lp=x_lp, lp_kind=kind)
        return x2, x_lp


class _TransformerBase(nn.Module):
    def __init__(self, d_model, d_hidden, n_layers, n_heads, drop_ratio, pe):
        super().__init__()
        if pe:
            raise NotImplementedError('pe=True is not implemented in the reference either '
                                      '(code/transformer_code.py:110-113)')
        self.encoder = _Stack(d_model, d_hidden, n_layers, n_heads, drop_ratio)
        self.compute = 'fp32x'
        self._exec = EncoderExecutor(self.encoder, d_model, n_heads, drop_ratio)

    def set_compute(self, mode):
        if mode not in COMPUTE_MODES:
            raise ValueError(f'compute must be one of {COMPUTE_MODES}')
        self.compute = mode
        return self


class Transformer(_TransformerBase):
    """Plain encoder stack (use_rel=False, the yml default): code/transformer_code.py:244-260."""

    def __init__(self, d_model, n_vocab_src, vocab_trg, d_hidden=2048, n_layers=6, n_heads=8,
                 drop_ratio=0.1, pe=False):
        super().__init__(d_model, d_hidden, n_layers, n_heads, drop_ratio, pe)

    def forward(self, x):
        return self._exec.run(x, None, self.compute, self.training)


class RelTransformer(_TransformerBase):
    """Relative-position-bias encoder stack: code/transformer_code.py:263-279."""

    def __init__(self, d_model, n_vocab_src, vocab_trg, d_hidden=2048, n_layers=6, n_heads=8,
                 drop_ratio=0.1, pe=False, d_pe=None):
        super().__init__(d_model, d_hidden, n_layers, n_heads, drop_ratio, pe)

    def forward(self, x, x_pe):
        return self._exec.run(x, x_pe, self.compute, self.training)
