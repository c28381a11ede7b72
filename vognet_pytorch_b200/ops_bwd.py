"""Tensor-level wrappers over the BACKWARD entry points of the C ABI (include/vog_b200.h, section "training step").
Same rules as ``ops``: CUDA tensors only, raw ``data_ptr()``s, kernels enqueued on torch's current stream, no fallback."""
import ctypes

import torch

from . import _lib
from .ops import BIAS_DENSE, BIAS_NONE, BIAS_RANK1, LP_BF16, LP_NONE, LP_TF32, _LP_DTYPE, _ptr, _req, _rowmajor2d, _stream

_KIND_OF = {torch.float32: 0, torch.bfloat16: 1}


def _strides2d(t, name):
    if t.dim() != 2 or not t.is_cuda or t.dtype != torch.float32:
        raise TypeError(f'{name}: expected a 2-D CUDA float32 tensor')
    s0, s1 = t.stride()
    if t.shape[1] == 1:
        s1 = 1
    if t.shape[0] == 1:
        s0 = max(s0, 1) if s1 == 1 else 1
    if s0 != 1 and s1 != 1:
        raise ValueError(f'{name}: one of the two strides must be 1, got {t.stride()}')
    return s0, s1


def sgemm(a, b, bias=None, relu=False, out=None, accumulate=False):
    """out[M,N] (+)= (relu?)(a[M,K] @ b[K,N] + bias); a and b are arbitrary 2-D VIEWS with one unit stride each
    (``w.t()`` and column slices are fine) - exact fp32 (vog_sgemm_strided)."""
    M, K = a.shape
    K2, N = b.shape
    if K != K2:
        raise ValueError(f'sgemm: a is {tuple(a.shape)} but b is {tuple(b.shape)}')
    sam, sak = _strides2d(a, 'a')
    sbk, sbn = _strides2d(b, 'b')
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
        if accumulate:
            out.zero_()
    _req(out, torch.float32, 'out', 2)
    if bias is not None:
        _req(bias, torch.float32, 'bias', 1)
    _lib.check(_lib.lib().vog_sgemm_strided(_ptr(a), sam, sak, _ptr(b), sbk, sbn, _ptr(bias), _ptr(out),
                                            _rowmajor2d(out, 'out'), M, N, K, int(relu), int(accumulate), _stream()),
               'vog_sgemm_strided')
    return out


def colsum_acc(x, out):
    """out[n] += sum_m x[m,n]"""
    _req(x, torch.float32, 'x', 2), _req(out, torch.float32, 'out', 1)
    _lib.check(_lib.lib().vog_colsum_acc(_ptr(x), _rowmajor2d(x, 'x'), _ptr(out), x.shape[0], x.shape[1], _stream()),
               'vog_colsum_acc')
    return out


def relu_bwd(dy, act, dbias=None, out=None, lp_kind=LP_NONE, inplace=False, want_f32=True):
    """g = dy * [act > 0] -> (fp32 g or None, low-precision g or None); dbias += column sums of g."""
    _req(dy, torch.float32, 'dy', 2)
    if act.dtype not in _KIND_OF or act.shape != dy.shape:
        raise TypeError('relu_bwd: act must be fp32 / bf16 with the shape of dy')
    M, N = dy.shape
    if inplace:
        out = dy
    out_lp = torch.empty(M, N, device=dy.device, dtype=_LP_DTYPE[lp_kind]) if lp_kind != LP_NONE else None
    if out is None and (want_f32 or out_lp is None):
        out = torch.empty(M, N, device=dy.device, dtype=torch.float32)
    _lib.check(_lib.lib().vog_relu_bwd(_ptr(dy), _rowmajor2d(dy, 'dy'), _ptr(act), _rowmajor2d(act, 'act'), _KIND_OF[act.dtype],
                                       _ptr(out), _rowmajor2d(out, 'out') if out is not None else 0, _ptr(out_lp), N, lp_kind,
                                       _ptr(dbias), M, N, _stream()), 'vog_relu_bwd')
    return out, out_lp


def layernorm_bwd(dy, x, gamma, dgamma, dbeta, dxsum=None, eps=1e-5, want_f32=True, lp_kind=LP_NONE):
    """-> (dx fp32 or None, dx low precision or None); dgamma / dbeta / dxsum accumulated."""
    _req(dy, torch.float32, 'dy', 2), _req(x, torch.float32, 'x', 2)
    M, d = x.shape
    dx = torch.empty(M, d, device=x.device, dtype=torch.float32) if want_f32 else None
    dx_lp = torch.empty(M, d, device=x.device, dtype=_LP_DTYPE[lp_kind]) if lp_kind != LP_NONE else None
    _lib.check(_lib.lib().vog_layernorm_bwd(_ptr(dy), _rowmajor2d(dy, 'dy'), _ptr(x), _rowmajor2d(x, 'x'), _ptr(gamma),
                                            _ptr(dx), d, _ptr(dx_lp), d, lp_kind, _ptr(dgamma), _ptr(dbeta), _ptr(dxsum),
                                            M, d, float(eps), _stream()), 'vog_layernorm_bwd')
    return dx, dx_lp


def attn_bwd_f32(q, k, v, out, dout, lse, Bt, N, head_dims, inv_scale, bias_mode=BIAS_NONE, a=None, nbox=0, bpe=None,
                 dense=None, da=None, dbpe=None, want_ddense=False, drop_p=0.0, seed=0):
    """q,k,v [Bt*N, ld] views (heads = column chunks), out/dout [Bt*N, d] -> dqkv [Bt*N, 3d] (dq | dk | dv) and, for a
    dense bias, its gradient [Bt,N,N,H]; da / dbpe accumulated for the rank-1 bias."""
    H, d = len(head_dims), sum(head_dims)
    ld = _rowmajor2d(q, 'q')
    if _rowmajor2d(k, 'k') != ld or _rowmajor2d(v, 'v') != ld:
        raise ValueError('attn_bwd_f32: q, k, v must share one leading dimension')
    dqkv = torch.empty(Bt * N, 3 * d, device=q.device, dtype=torch.float32)
    delta = torch.empty(Bt * H * N, device=q.device, dtype=torch.float32)
    ddense = torch.empty_like(dense) if (want_ddense and dense is not None) else None
    offs = [sum(head_dims[:h]) for h in range(H)]
    off_arr = (ctypes.c_int * H)(*offs)
    dh_arr = (ctypes.c_int * H)(*head_dims)
    _lib.check(_lib.lib().vog_attn_bwd_f32(_ptr(q), _ptr(k), _ptr(v), ld, _ptr(out), _rowmajor2d(out, 'out'), _ptr(dout),
                                           _rowmajor2d(dout, 'dout'), _ptr(lse), _ptr(delta), _ptr(dqkv[:, :d]),
                                           _ptr(dqkv[:, d:2 * d]), _ptr(dqkv[:, 2 * d:]), 3 * d, Bt, N, H, off_arr, dh_arr,
                                           float(inv_scale), bias_mode, _ptr(a), nbox, _ptr(bpe), _ptr(dense), _ptr(da),
                                           _ptr(dbpe), _ptr(ddense), float(drop_p), int(seed), _stream()), 'vog_attn_bwd_f32')
    return dqkv, ddense


def pe_project_bwd(props, da, dW, vid_w, vid_h, fdiv):
    _req(props, torch.float32, 'props', 2), _req(da, torch.float32, 'da', 2), _req(dW, torch.float32, 'dW', 2)
    _lib.check(_lib.lib().vog_pe_project_bwd(_ptr(props), _rowmajor2d(props, 'props'), _ptr(da.contiguous()), _ptr(dW),
                                             props.shape[0], dW.shape[0], float(vid_w), float(vid_h), float(fdiv),
                                             _stream()), 'vog_pe_project_bwd')
    return dW


def xmul_bwd(dtok, dlang, B, nfrm, nsrl, nppf2, dv):
    """dtok [B*nfrm*nsrl*nppf2, dv+dl] -> dvis [B*nfrm*nppf2, dv]; dlang [B*nsrl, dl] accumulated."""
    _req(dtok, torch.float32, 'dtok', 2)
    if not dtok.is_contiguous():
        raise ValueError('xmul_bwd: dtok must be contiguous')
    dl = dtok.shape[1] - dv
    dvis = torch.empty(B * nfrm * nppf2, dv, device=dtok.device, dtype=torch.float32)
    _lib.check(_lib.lib().vog_xmul_bwd(_ptr(dtok), _ptr(dvis), _ptr(dlang), B, nfrm, nsrl, nppf2, dv, dl, _stream()),
               'vog_xmul_bwd')
    return dvis


def seg_rep_bwd(dx, x, pe, se, nppf):
    nslots = dx.shape[0] // nppf
    dseg = torch.empty(nslots, se, device=dx.device, dtype=torch.float32)
    _lib.check(_lib.lib().vog_seg_rep_bwd(_ptr(dx), _ptr(x), _rowmajor2d(dx, 'dx'), pe, se, nppf, _ptr(dseg), nslots,
                                          _stream()), 'vog_seg_rep_bwd')
    return dseg


def lin2_bwd(dlogits, h, w2, dw2, db2, db1, nfrm, nsrl, nppf2, want_f32=True, lp_kind=LP_NONE):
    """dlogits [B,nsrl,P] contiguous, h [M,K] -> (dh fp32 or None, dh lp or None); dw2/db2/db1 accumulated."""
    _req(dlogits, torch.float32, 'dlogits')
    M, K = h.shape
    dh = torch.empty(M, K, device=h.device, dtype=torch.float32) if want_f32 else None
    dh_lp = torch.empty(M, K, device=h.device, dtype=_LP_DTYPE[lp_kind]) if lp_kind != LP_NONE else None
    _lib.check(_lib.lib().vog_lin2_bwd(_ptr(dlogits.contiguous()), _ptr(h), _rowmajor2d(h, 'h'), _KIND_OF[h.dtype],
                                       _ptr(w2.contiguous()), _ptr(dh), _ptr(dh_lp), lp_kind, _ptr(dw2), _ptr(db2), _ptr(db1),
                                       M, K, nfrm, nsrl, nppf2, _stream()), 'vog_lin2_bwd')
    return dh, dh_lp


def lang_gather_bwd(dcat, cap, T, Bq):
    """dcat [Bq*nsrl, 2D], cap [Bq,nsrl,2] -> dfull [T*Bq, D]"""
    _req(dcat, torch.float32, 'dcat', 2), _req(cap, torch.int64, 'cap', 3)
    D, nsrl = dcat.shape[1] // 2, cap.shape[1]
    dfull = torch.zeros(T * Bq, D, device=dcat.device, dtype=torch.float32)
    _lib.check(_lib.lib().vog_lang_gather_bwd(_ptr(dcat.contiguous()), D, _ptr(cap.contiguous()), T, Bq, nsrl, _ptr(dfull),
                                              _stream()), 'vog_lang_gather_bwd')
    return dfull


def lang_embed_bwd(words, mask, dx, pad_idx, lens, demb):
    _req(words, torch.int64, 'words', 2), _req(mask, torch.int64, 'mask', 2), _req(dx, torch.float32, 'dx', 2)
    Bq, T = mask.shape
    _lib.check(_lib.lib().vog_lang_embed_bwd(_ptr(words.contiguous()), words.shape[1], _ptr(mask.contiguous()), T,
                                             _ptr(dx.contiguous()), dx.shape[1], int(pad_idx), Bq, _ptr(lens), _ptr(demb),
                                             _stream()), 'vog_lang_embed_bwd')
    return demb


def lstm_hprev(hout, lens, T, Bq):
    _req(hout, torch.float32, 'hout', 2)
    H = hout.shape[1] // 2
    hp = torch.empty_like(hout)
    _lib.check(_lib.lib().vog_lstm_hprev(_ptr(hout.contiguous()), _ptr(lens), _ptr(hp), T, Bq, H, _stream()), 'vog_lstm_hprev')
    return hp


def lstm_scan(G, lens, T, Bq):
    H = G.shape[1] // 8
    acts = torch.empty(T * Bq, 2, 6, H, device=G.device, dtype=torch.float32)
    _lib.check(_lib.lib().vog_lstm_scan(_ptr(G), _ptr(lens), _ptr(acts), T, Bq, H, _stream()), 'vog_lstm_scan')
    return acts


def lstm_bwd_steps(dout, acts, whh_t, lens, T, Bq, whh=None):
    """dout [T*Bq, 2H] -> dG [T*Bq, 8H]; whh_t [2,H,4H] = the recurrent weights TRANSPOSED (whh.transpose(1,2));
    at most 8 sequences per call (callers chunk the batch)."""
    H = whh_t.shape[1]
    if whh_t.shape != (2, H, 4 * H) or not whh_t.is_contiguous():
        raise ValueError(f'lstm_bwd_steps: whh_t must be contiguous [2,H,4H], got {tuple(whh_t.shape)}')
    dG = torch.empty(T * Bq, 8 * H, device=dout.device, dtype=torch.float32)
    L = _lib.lib()
    nws = int(L.vog_lstm_bwd_workspace_bytes(Bq, H))
    ws = torch.empty(nws, device=dout.device, dtype=torch.uint8)
    if whh is not None and (whh.shape != (2, 4 * H, H) or not whh.is_contiguous() or whh.dtype != torch.float32):
        raise ValueError(f'lstm_bwd_steps: whh must be contiguous fp32 [2,4H,H], got {tuple(whh.shape)}')
    _lib.check(L.vog_lstm_bwd_steps(_ptr(dout.contiguous()), _ptr(acts), _ptr(whh_t), _ptr(whh), _ptr(lens), _ptr(dG), _ptr(ws), nws,
                                    T, Bq, H, _stream()), 'vog_lstm_bwd_steps')
    return dG


# ---------------------------------------------------------------------------------------------
# tensor-core training step ('bf16')
# ---------------------------------------------------------------------------------------------
def _bf16_2d(t, name):
    if t.dim() != 2 or not t.is_cuda or t.dtype != torch.bfloat16 or t.stride(1) != 1:
        raise TypeError(f'{name}: expected a 2-D CUDA bfloat16 tensor with unit column stride')
    return t.stride(0)


def tc_gemm_tn(a, b, out=None):
    """out[N1,N2] (fp32) += a[K,N1]^T @ b[K,N2]; a, b bf16 row-major views (vog_tc_gemm_tn).  out=None: fresh zeros."""
    lda, ldb = _bf16_2d(a, 'a'), _bf16_2d(b, 'b')
    K, N1 = a.shape
    if b.shape[0] != K:
        raise ValueError(f'tc_gemm_tn: a is {tuple(a.shape)} but b is {tuple(b.shape)}')
    N2 = b.shape[1]
    if out is None:
        out = torch.zeros(N1, N2, device=a.device, dtype=torch.float32)
    _req(out, torch.float32, 'out', 2)
    _lib.check(_lib.lib().vog_tc_gemm_tn(_ptr(a), lda, _ptr(b), ldb, K, N1, N2, _ptr(out), _rowmajor2d(out, 'out'), _stream()),
               'vog_tc_gemm_tn')
    return out


def tc_attn_fwd_train(q, k, v, N, head_dims, inv_scale, bias_mode=BIAS_NONE, a=None, nbox=0, bpe=None, drop_p=0.0, seed=0):
    """q,k,v [Bt,H,N,dhp] bf16 -> (out [Bt*N, H*dhp] bf16, lse [Bt,H,N] fp32 in the log2 domain)."""
    Bt, H, Nq, dhp = q.shape
    if Nq != N or k.shape != q.shape or v.shape != q.shape or len(head_dims) != H:
        raise ValueError('tc_attn_fwd_train: inconsistent shapes')
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _req(t, torch.bfloat16, n, 4)
        if not t.is_contiguous():
            raise ValueError(f'tc_attn_fwd_train: {n} must be contiguous')
    out = torch.empty(Bt * N, H * dhp, device=q.device, dtype=torch.bfloat16)
    lse = torch.empty(Bt, H, N, device=q.device, dtype=torch.float32)
    dh_arr = (ctypes.c_int * H)(*head_dims)
    L = _lib.lib()
    ws, ws_bytes = None, 0
    if bias_mode == BIAS_RANK1:
        _req(a, torch.float32, 'a', 2), _req(bpe, torch.float32, 'bpe', 1)
        if a.shape != (Bt * nbox, H) or not a.is_contiguous():
            raise ValueError(f'tc_attn_fwd_train: a must be contiguous [{Bt * nbox},{H}], got {tuple(a.shape)}')
        ws_bytes = L.vog_tc_attn_workspace_bytes(Bt, N, H)
        ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
    elif bias_mode != BIAS_NONE:
        raise ValueError('tc_attn_fwd_train: rank-1 bias or none (a dense x_pe trains in the fp32x mode)')
    _lib.check(L.vog_tc_attn_fwd_train(_ptr(q), _ptr(k), _ptr(v), Bt, N, H, dhp, dh_arr, float(inv_scale), bias_mode,
                                       _ptr(a), nbox, _ptr(bpe), _ptr(out), H * dhp, LP_BF16, _ptr(ws), ws_bytes,
                                       _ptr(lse), float(drop_p), int(seed), _stream()), 'vog_tc_attn_fwd_train')
    return out, lse


def tc_attn_bwd(q, k, v, out, dout, lse, N, head_dims, inv_scale, bias_mode=BIAS_NONE, a=None, nbox=0, bpe=None, da=None,
                dbpe=None, drop_p=0.0, seed=0):
    """-> dqkv [Bt*N, 3*H*dhp] bf16 (dQ | dK | dV, padded head slots); da [Bt*nbox,H] / dbpe [H] accumulated."""
    Bt, H, Nq, dhp = q.shape
    ldo, lddo = _bf16_2d(out, 'out'), _bf16_2d(dout, 'dout')
    dqkv = torch.empty(Bt * N, 3 * H * dhp, device=q.device, dtype=torch.bfloat16)
    L = _lib.lib()
    ws_bytes = L.vog_tc_attn_bwd_workspace_bytes(Bt, N, H)
    ws = torch.empty(ws_bytes + 256, device=q.device, dtype=torch.uint8)
    off = (-ws.data_ptr()) % 256
    dh_arr = (ctypes.c_int * H)(*head_dims)
    if bias_mode == BIAS_RANK1:
        _req(da, torch.float32, 'da', 2), _req(dbpe, torch.float32, 'dbpe', 1)
        if da.shape != a.shape or not da.is_contiguous():
            raise ValueError('tc_attn_bwd: da must be contiguous with the shape of a')
    _lib.check(L.vog_tc_attn_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(out), ldo, _ptr(dout), lddo, _ptr(lse), Bt, N, H, dhp,
                                 dh_arr, float(inv_scale), bias_mode, _ptr(a), nbox, _ptr(bpe), _ptr(dqkv), 3 * H * dhp,
                                 _ptr(da), _ptr(dbpe), ws.data_ptr() + off, ws_bytes, float(drop_p), int(seed), _stream()),
               'vog_tc_attn_bwd')
    return dqkv


def dropout(x, p, seed, stream_id, residual=None, out=None, lp_kind=LP_NONE, want_f32=True):
    """x [M,N] fp32 -> (x * keep/(1-p) + residual as fp32 or None, low-precision copy or None); the mask is a function of
    (seed, stream_id, row, column) (vog_dropout).  The backward is the same call on the output gradient."""
    _req(x, torch.float32, 'x', 2)
    M, N = x.shape
    if out is None and want_f32:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    out_lp = torch.empty(M, N, device=x.device, dtype=_LP_DTYPE[lp_kind]) if lp_kind != LP_NONE else None
    ldr = 0
    if residual is not None:
        _req(residual, torch.float32, 'residual', 2)
        ldr = _rowmajor2d(residual, 'residual')
    _lib.check(_lib.lib().vog_dropout(_ptr(x), _rowmajor2d(x, 'x'), _ptr(residual), ldr, _ptr(out),
                                      _rowmajor2d(out, 'out') if out is not None else 0, _ptr(out_lp), N, lp_kind, M, N,
                                      float(p), int(seed), int(stream_id), _stream()), 'vog_dropout')
    return out, out_lp
