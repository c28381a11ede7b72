"""Tensor-level wrappers over the C ABI (include/vog_b200.h).

PyTorch is used here for device memory and streams only: every function checks its operands,
takes raw ``data_ptr()``s and enqueues the CUDA kernel on torch's current stream.  CPU tensors are
rejected - there is no fallback implementation.
"""
import ctypes
import math

import torch

from . import _lib

BIAS_NONE, BIAS_RANK1, BIAS_DENSE, BIAS_RANK1_EXPANDED = 0, 1, 2, 3
LP_NONE, LP_BF16, LP_TF32 = 0, 1, 2


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype, name, dims=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'{name}: expected a CUDA tensor (vognet_pytorch_b200 has no CPU path)')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if dims is not None and t.dim() != dims:
        raise ValueError(f'{name}: expected {dims} dims, got shape {tuple(t.shape)}')


def _rowmajor2d(t, name):
    """(ld) of a 2-D view whose rows are contiguous (column slices of a wider matrix are fine)."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f'{name}: need a row-major 2-D tensor, got shape {tuple(t.shape)} strides {t.stride()}')
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def chunk_sizes(d, n_heads):
    """torch.chunk split of the model dimension (code/transformer_code.py:66-67,182-183)."""
    c = -(-d // n_heads)
    out, left = [], d
    while left > 0:
        out.append(min(c, left))
        left -= c
    return out


# ---------------------------------------------------------------------------------------------
def sgemm_nt(a, w, bias=None, residual=None, relu=False, out=None):
    """out[M,N] = (relu?)(a[M,K] @ w[N,K]^T + bias) + residual        exact fp32"""
    _req(a, torch.float32, 'a', 2), _req(w, torch.float32, 'w', 2)
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f'sgemm_nt: a is [{M},{K}] but w is {tuple(w.shape)}')
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    _req(out, torch.float32, 'out', 2)
    if bias is not None:
        _req(bias, torch.float32, 'bias', 1)
    ldr = 0
    if residual is not None:
        _req(residual, torch.float32, 'residual', 2)
        ldr = _rowmajor2d(residual, 'residual')
    L = _lib.lib()
    _lib.check(L.vog_sgemm_nt(_ptr(a), _rowmajor2d(a, 'a'), _ptr(w), _rowmajor2d(w, 'w'), _ptr(bias),
                              _ptr(residual), ldr, _ptr(out), _rowmajor2d(out, 'out'), M, N, K,
                              int(relu), _stream()), 'vog_sgemm_nt')
    return out


def attn_fwd_f32(q, k, v, Bt, N, head_dims, inv_scale, out=None, bias_mode=BIAS_NONE, a=None,
                 nbox=0, bpe=None, dense=None, lse=None, drop_p=0.0, seed=0):
    """q,k,v: [Bt*N, d] (row-major views, same ld); heads are consecutive column chunks.  ``lse`` (optional
    [Bt,H,N] fp32) receives the log-sum-exp of the scaled scores for the backward; ``drop_p`` / ``seed``: training-mode
    dropout on the probabilities."""
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _req(t, torch.float32, n, 2)
    ld = _rowmajor2d(q, 'q')
    if _rowmajor2d(k, 'k') != ld or _rowmajor2d(v, 'v') != ld:
        raise ValueError('attn_fwd_f32: q, k, v must share one leading dimension')
    H = len(head_dims)
    d = sum(head_dims)
    if out is None:
        out = torch.empty(Bt * N, d, device=q.device, dtype=torch.float32)
    offs = [sum(head_dims[:h]) for h in range(H)]
    off_arr = (ctypes.c_int * H)(*offs)
    dh_arr = (ctypes.c_int * H)(*head_dims)
    if bias_mode == BIAS_RANK1:
        _req(a, torch.float32, 'a', 2)
        if a.shape != (Bt * nbox, H) or not a.is_contiguous():
            raise ValueError(f'attn_fwd_f32: a must be contiguous [{Bt * nbox},{H}], got {tuple(a.shape)}')
        _req(bpe, torch.float32, 'bpe', 1)
    if bias_mode == BIAS_DENSE:
        _req(dense, torch.float32, 'dense', 4)
        if tuple(dense.shape) != (Bt, N, N, H) or not dense.is_contiguous():
            raise ValueError(f'attn_fwd_f32: dense bias must be contiguous [{Bt},{N},{N},{H}]')
    L = _lib.lib()
    _lib.check(L.vog_attn_fwd_f32(_ptr(q), _ptr(k), _ptr(v), ld, _ptr(out), _rowmajor2d(out, 'out'),
                                  Bt, N, H, off_arr, dh_arr, float(inv_scale), bias_mode, _ptr(a),
                                  nbox, _ptr(bpe), _ptr(dense), _ptr(lse), float(drop_p), int(seed), _stream()),
               'vog_attn_fwd_f32')
    return out


def add_layernorm(x, residual, weight, bias, eps=1e-5, out=None, out_lp=None, lp_kind=LP_NONE):
    _req(x, torch.float32, 'x', 2)
    M, d = x.shape
    if out is None and out_lp is None:
        out = torch.empty(M, d, device=x.device, dtype=torch.float32)
    ldr = 0
    if residual is not None:
        _req(residual, torch.float32, 'residual', 2)
        ldr = _rowmajor2d(residual, 'residual')
    L = _lib.lib()
    _lib.check(L.vog_add_layernorm(_ptr(x), _rowmajor2d(x, 'x'), _ptr(residual), ldr, _ptr(weight),
                                   _ptr(bias), _ptr(out), _rowmajor2d(out, 'out') if out is not None else 0,
                                   _ptr(out_lp), _rowmajor2d(out_lp, 'out_lp') if out_lp is not None else 0,
                                   lp_kind, M, d, float(eps), _stream()), 'vog_add_layernorm')
    return out if out is not None else out_lp


def pe_project(props, W, vid_w, vid_h, fdiv, scale=1.0):
    """props [rows, >=5] (row-major view) -> a [rows, H] = scale * W . normalised(props[:, :5])"""
    _req(props, torch.float32, 'props', 2), _req(W, torch.float32, 'W', 2)
    rows, H = props.shape[0], W.shape[0]
    a = torch.empty(rows, H, device=props.device, dtype=torch.float32)
    L = _lib.lib()
    _lib.check(L.vog_pe_project(_ptr(props), _rowmajor2d(props, 'props'), _ptr(W.contiguous()), _ptr(a),
                                rows, H, float(vid_w), float(vid_h), float(fdiv), float(scale),
                                _stream()), 'vog_pe_project')
    return a


def pe_project_expand(props, W, vid_w, vid_h, fdiv, Bt, N, nbox, inv_scale, scale=1.0):
    """pe_project plus, in the same launch, the per-key bias factors of the attention call (Bt sequences of N tokens,
    bias period nbox, scale inv_scale) -> (a [rows, H], key_factors): pass key_factors to tc_attn_fwd(ak=...)."""
    _req(props, torch.float32, 'props', 2), _req(W, torch.float32, 'W', 2)
    rows, H = props.shape[0], W.shape[0]
    if rows != Bt * nbox:
        raise ValueError(f'pe_project_expand: {rows} proposal rows != Bt*nbox = {Bt}*{nbox}')
    L = _lib.lib()
    a = torch.empty(rows, H, device=props.device, dtype=torch.float32)
    nb = int(L.vog_tc_attn_workspace_bytes(Bt, N, H))
    ak = torch.empty(nb, device=props.device, dtype=torch.uint8)
    _lib.check(L.vog_pe_project_expand(_ptr(props), _rowmajor2d(props, 'props'), _ptr(W.contiguous()), _ptr(a), rows, H,
                                       float(vid_w), float(vid_h), float(fdiv), float(scale), Bt, N, nbox,
                                       float(inv_scale), _ptr(ak), nb, _stream()), 'vog_pe_project_expand')
    return a, ak


def select_fwd(scores, props, ncmp, nfrm, nppf, spat):
    """scores [B,nsrl,P], props [B,P,pdim] -> boxes, scores, indexs (see vog_select_fwd)."""
    _req(scores, torch.float32, 'scores', 3), _req(props, torch.float32, 'props', 3)
    B, nsrl, Pn = scores.shape
    pdim = props.shape[-1]
    if Pn != ncmp * nfrm * nppf or props.shape[1] != Pn:
        raise ValueError(f'select_fwd: P={Pn} != ncmp*nfrm*nppf={ncmp * nfrm * nppf}')
    scores, props = scores.contiguous(), props.contiguous()
    # the three results are views of ONE allocation (int64 first: alignment), so a caller that wants them on the host
    # can move them with a single copy (runtime.PredictionFetcher)
    n_ix, n_bx, n_sc = B * nsrl * nfrm * 8, B * nsrl * ncmp * nfrm * pdim * 4, B * nsrl * ncmp * nfrm * 4
    flat = torch.empty(n_ix + n_bx + n_sc, device=scores.device, dtype=torch.uint8)
    ix = flat[:n_ix].view(torch.int64).view(B, nsrl, nfrm)
    boxes = flat[n_ix:n_ix + n_bx].view(torch.float32).view(B, nsrl, ncmp, nfrm, pdim)
    sc = flat[n_ix + n_bx:].view(torch.float32).view(B, nsrl, ncmp, nfrm)
    L = _lib.lib()
    _lib.check(L.vog_select_fwd(_ptr(scores), _ptr(props), pdim, _ptr(boxes), _ptr(sc), _ptr(ix),
                                B, nsrl, ncmp, nfrm, nppf, int(spat), _stream()), 'vog_select_fwd')
    return boxes, sc, ix


def concat_videos(feat, seg, props, conc_type, nfrm, nppf, vid_shift=None):
    """Per-video tensors feat [B,ncmp,nfrm*nppf,D], seg [B,ncmp,nfrm,Ds], props [B,ncmp,nfrm*nppf,pdim] -> the
    concatenated (feat [B,P,D], seg [B,ncmp*nfrm,Ds], props [B,P,pdim]) of conc_type 'spat' / 'temp' (see
    vog_concat_videos).  TEMP keeps the row order: its feat / seg results are views of the inputs."""
    if conc_type not in ('spat', 'temp'):
        raise ValueError("concat_videos: conc_type must be 'spat' or 'temp'")
    _req(feat, torch.float32, 'feat', 4), _req(seg, torch.float32, 'seg', 4), _req(props, torch.float32, 'props', 4)
    B, ncmp, P1, D = feat.shape
    if P1 != nfrm * nppf or tuple(seg.shape[:3]) != (B, ncmp, nfrm) or tuple(props.shape[:3]) != (B, ncmp, P1):
        raise ValueError(f'concat_videos: feat {tuple(feat.shape)} / seg {tuple(seg.shape)} / props {tuple(props.shape)} '
                         f'do not describe [B,ncmp,{nfrm}*{nppf}]')
    spat = conc_type == 'spat'
    shift = float(vid_shift if vid_shift is not None else (720.0 if spat else 10.0))
    feat, seg, props = feat.contiguous(), seg.contiguous(), props.contiguous()
    props_out = torch.empty(B, ncmp * P1, props.shape[-1], device=props.device, dtype=torch.float32)
    if spat:
        feat_out = torch.empty(B, ncmp * P1, D, device=feat.device, dtype=torch.float32)
        seg_out = torch.empty(B, ncmp * nfrm, seg.shape[-1], device=seg.device, dtype=torch.float32)
    else:
        feat_out, seg_out = None, None
    _lib.check(_lib.lib().vog_concat_videos(_ptr(feat), D, _ptr(seg), seg.shape[-1], _ptr(props), props.shape[-1],
                                            _ptr(feat_out), _ptr(seg_out), _ptr(props_out), B, ncmp, nfrm, nppf,
                                            int(spat), shift, _stream()), 'vog_concat_videos')
    if not spat:
        feat_out, seg_out = feat.view(B, ncmp * P1, D), seg.view(B, ncmp * nfrm, seg.shape[-1])
    return feat_out, seg_out, props_out


def verb_loss_fwd(vidf, verb_cmp, vcc_msk, loss_lambda=1.0):
    """vidf [n] f32, verb_cmp [n] int64, vcc_msk [n,m] int64 -> loss [1] (see vog_verb_loss_fwd)."""
    _req(vidf, torch.float32, 'vidf', 1), _req(verb_cmp, torch.int64, 'verb_cmp', 1), _req(vcc_msk, torch.int64, 'vcc_msk', 2)
    n, m = vcc_msk.shape
    if vidf.shape[0] != n or verb_cmp.shape[0] != n:
        raise ValueError('verb_loss_fwd: inconsistent shapes')
    loss = torch.empty(1, device=vidf.device, dtype=torch.float32)
    _lib.check(_lib.lib().vog_verb_loss_fwd(_ptr(vidf.contiguous()), _ptr(verb_cmp.contiguous()), _ptr(vcc_msk.contiguous()),
                                            n, m, float(loss_lambda), _ptr(loss), _stream()), 'vog_verb_loss_fwd')
    return loss


def select_sep_fwd(scores, props, fin_scores, nfrm, nppf):
    """scores [B,ncmp,nsrl,nfrm*nppf], props [B,ncmp,nfrm*nppf,pdim], fin_scores [B,ncmp] -> boxes, scores, indexs
    (see vog_select_sep_fwd)."""
    _req(scores, torch.float32, 'scores', 4), _req(props, torch.float32, 'props', 4)
    _req(fin_scores, torch.float32, 'fin_scores', 2)
    B, ncmp, nsrl, P1 = scores.shape
    pdim = props.shape[-1]
    if P1 != nfrm * nppf or tuple(props.shape[:3]) != (B, ncmp, P1) or tuple(fin_scores.shape) != (B, ncmp):
        raise ValueError(f'select_sep_fwd: scores {tuple(scores.shape)} / props {tuple(props.shape)} / fin_scores '
                         f'{tuple(fin_scores.shape)} do not describe [B,ncmp,nsrl,{nfrm}*{nppf}]')
    scores, props, fin_scores = scores.contiguous(), props.contiguous(), fin_scores.contiguous()
    boxes = torch.empty(B, nsrl, ncmp, nfrm, pdim, device=scores.device, dtype=torch.float32)
    sc = torch.empty(B, nsrl, ncmp, nfrm, device=scores.device, dtype=torch.float32)
    ix = torch.empty(B, nsrl, nfrm, device=scores.device, dtype=torch.int64)
    L = _lib.lib()
    _lib.check(L.vog_select_sep_fwd(_ptr(scores), _ptr(props), pdim, _ptr(fin_scores), _ptr(boxes), _ptr(sc), _ptr(ix),
                                    B, nsrl, ncmp, nfrm, nppf, _stream()), 'vog_select_sep_fwd')
    return boxes, sc, ix


def sep_fin_scores(logits, vidf, srl_msk, verb_ind, cmp_msk):
    """logits [Bq,nsrl,P1], vidf [Bq], srl_msk [Bq,nsrl] / verb_ind [Bq] / cmp_msk [Bq] int64 -> fin_loss [Bq,nsrl],
    fin_eval [Bq] (see vog_sep_fin_scores)."""
    _req(logits, torch.float32, 'logits', 3), _req(vidf, torch.float32, 'vidf', 1)
    Bq, nsrl, P1 = logits.shape
    for t, nm, shp in ((srl_msk, 'srl_msk', (Bq, nsrl)), (verb_ind, 'verb_ind', (Bq,)), (cmp_msk, 'cmp_msk', (Bq,))):
        if t.dtype != torch.int64 or tuple(t.shape) != shp or not t.is_cuda:
            raise ValueError(f'sep_fin_scores: {nm} must be a CUDA int64 tensor of shape {shp}')
    logits, vidf = logits.contiguous(), vidf.contiguous()
    srl_msk, verb_ind, cmp_msk = srl_msk.contiguous(), verb_ind.contiguous(), cmp_msk.contiguous()
    fin_loss = torch.empty(Bq, nsrl, device=logits.device, dtype=torch.float32)
    fin_eval = torch.empty(Bq, device=logits.device, dtype=torch.float32)
    L = _lib.lib()
    _lib.check(L.vog_sep_fin_scores(_ptr(logits), _ptr(vidf), _ptr(srl_msk), _ptr(verb_ind), _ptr(cmp_msk),
                                    _ptr(fin_loss), _ptr(fin_eval), Bq, nsrl, P1, _stream()), 'vog_sep_fin_scores')
    return fin_loss, fin_eval


def inv_sqrt(d_model):
    return 1.0 / math.sqrt(d_model)


# ---------------------------------------------------------------------------------------------
# tensor-core path
# ---------------------------------------------------------------------------------------------
_LP_DTYPE = {LP_BF16: torch.bfloat16, LP_TF32: torch.float32}


def cast_lp(src, kind, out=None):
    """fp32 [rows, cols] (row-major view) -> bf16 or tf32-rounded fp32 copy."""
    _req(src, torch.float32, 'src', 2)
    rows, cols = src.shape
    if out is None:
        out = torch.empty(rows, cols, device=src.device, dtype=_LP_DTYPE[kind])
    L = _lib.lib()
    _lib.check(L.vog_cast_lp(_ptr(src), _rowmajor2d(src, 'src'), _ptr(out), _rowmajor2d(out, 'out'),
                             rows, cols, kind, _stream()), 'vog_cast_lp')
    return out


def _is_tf32(a, w):
    if a.dtype != w.dtype or a.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError(f'tc_gemm operands must both be bf16 or both tf32-rounded fp32, got {a.dtype}/{w.dtype}')
    return int(a.dtype == torch.float32)


_SMS = {}


def _num_sms(device):
    idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if idx not in _SMS:
        _SMS[idx] = torch.cuda.get_device_properties(idx).multi_processor_count
    return _SMS[idx]


def pick_bn(N, M=None, sms=148):
    """Column-tile width of the tcgen05 GEMM: the largest divisor of N in {256,192,128,64,32} that still
    yields enough 128 x BN tiles to occupy most SMs (small-M problems get narrow tiles instead of split-K
    partials + a second reduction launch); plain divisibility when M is unknown."""
    cands = [bn for bn in (256, 192, 128, 64, 32) if N % bn == 0]
    if not cands:
        return 128 if N > 64 else (64 if N > 32 else 32)
    if M is None:
        return cands[0]
    mblk = -(-M // 128)
    for bn in cands:
        if mblk * (N // bn) >= 0.6 * sms:
            return bn
    return cands[-1]


def tc_gemm(a, w, bias=None, residual=None, relu=False, out_f32=None, out_lp=None, lp_kind=LP_NONE,
            rep=1, BN=None, want_f32=True):
    """tcgen05 GEMM: (relu?)(a @ w^T + bias) + residual -> fp32 and/or low-precision outputs."""
    if not (isinstance(a, torch.Tensor) and a.is_cuda and w.is_cuda):
        raise RuntimeError('tc_gemm: expected CUDA tensors (vognet_pytorch_b200 has no CPU path)')
    tf32 = _is_tf32(a, w)
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise ValueError(f'tc_gemm: a is [{M},{K}] but w is {tuple(w.shape)}')
    if out_f32 is None and want_f32:
        out_f32 = torch.empty(M * rep, N, device=a.device, dtype=torch.float32)
    if out_lp is None and lp_kind != LP_NONE:
        out_lp = torch.empty(M * rep, N, device=a.device, dtype=_LP_DTYPE[lp_kind])
    if out_lp is not None and lp_kind == LP_NONE:
        lp_kind = LP_BF16 if out_lp.dtype == torch.bfloat16 else LP_TF32
    ldr = 0
    if residual is not None:
        _req(residual, torch.float32, 'residual', 2)
        ldr = _rowmajor2d(residual, 'residual')
    if bias is not None:
        _req(bias, torch.float32, 'bias', 1)
    L = _lib.lib()
    BN = BN or pick_bn(N, M, _num_sms(a.device))
    ws, ws_bytes = None, L.vog_tc_gemm_workspace_bytes(M, N, K, tf32, BN)
    if ws_bytes:
        ws = torch.empty(ws_bytes, device=a.device, dtype=torch.uint8)
    _lib.check(L.vog_tc_gemm(_ptr(a), _rowmajor2d(a, 'a'), _ptr(w), _rowmajor2d(w, 'w'), M, N, K, tf32,
                             BN, _ptr(bias), int(relu), _ptr(residual), ldr,
                             _ptr(out_f32), _rowmajor2d(out_f32, 'out_f32') if out_f32 is not None else 0,
                             _ptr(out_lp), _rowmajor2d(out_lp, 'out_lp') if out_lp is not None else 0,
                             lp_kind, rep, _ptr(ws), ws_bytes, _stream()), 'vog_tc_gemm')
    return out_f32, out_lp


def round_up(a, b):
    return -(-a // b) * b


def pack_weights(wq, wk, wv, wo, head_dims, dhp, kind):
    """fp32 [d,d] projection weights of one attention block -> (wqkv [3*H*dhp, d], wo_p [d, H*dhp]) with per-head zero
    padding, in bf16 / tf32-rounded fp32 (vog_pack_weights)."""
    ws = [w.detach().float().contiguous() for w in (wq, wk, wv, wo)]
    d, H = ws[0].shape[1], len(head_dims)
    for w in ws:
        _req(w, torch.float32, 'weight', 2)
        if tuple(w.shape) != (d, d):
            raise ValueError(f'pack_weights: expected [{d},{d}] weights, got {tuple(w.shape)}')
    wqkv = torch.empty(3 * H * dhp, d, device=ws[0].device, dtype=_LP_DTYPE[kind])
    wo_p = torch.empty(d, H * dhp, device=ws[0].device, dtype=_LP_DTYPE[kind])
    dh_arr = (ctypes.c_int * H)(*head_dims)
    _lib.check(_lib.lib().vog_pack_weights(_ptr(ws[0]), _ptr(ws[1]), _ptr(ws[2]), _ptr(ws[3]), d, H, dh_arr, dhp, kind,
                                           _ptr(wqkv), _ptr(wo_p), _stream()), 'vog_pack_weights')
    return wqkv, wo_p


def tc_gemm_qkv(a, wqkv, Bt, N, n_heads, dhp):
    """-> q,k,v [Bt,H,N,dhp] bf16."""
    tf32 = _is_tf32(a, wqkv)
    M, K = a.shape
    assert M == Bt * N and wqkv.shape == (3 * n_heads * dhp, K)
    q = torch.empty(Bt, n_heads, N, dhp, device=a.device, dtype=torch.bfloat16)
    k, v = torch.empty_like(q), torch.empty_like(q)
    L = _lib.lib()
    _lib.check(L.vog_tc_gemm_qkv(_ptr(a), _rowmajor2d(a, 'a'), _ptr(wqkv), _rowmajor2d(wqkv, 'wqkv'), M, K,
                                 tf32, n_heads, dhp, N, _ptr(q), _ptr(k), _ptr(v), _stream()),
               'vog_tc_gemm_qkv')
    return q, k, v


def tc_gemm_qkv_factored(vis_lp, wqkv_vis, lq, Bt, nfrm, nsrl, nppf2, n_heads, dhp):
    """Factorised QKV projection (see vog_tc_gemm_qkv_factored): vis_lp [Bt*nppf2, dv] low precision,
    wqkv_vis [3*H*dhp, dv] (a column-slice view of the packed weight is fine), lq [B*nsrl, 3*H*dhp] fp32
    -> q,k,v [Bt,H,nsrl*nppf2,dhp] bf16."""
    tf32 = _is_tf32(vis_lp, wqkv_vis)
    M, K = vis_lp.shape
    _req(lq, torch.float32, 'lq', 2)
    N = nsrl * nppf2
    if M != Bt * nppf2 or wqkv_vis.shape != (3 * n_heads * dhp, K) or Bt % nfrm != 0 or \
            lq.shape != ((Bt // nfrm) * nsrl, 3 * n_heads * dhp):
        raise ValueError('tc_gemm_qkv_factored: inconsistent shapes')
    q = torch.empty(Bt, n_heads, N, dhp, device=vis_lp.device, dtype=torch.bfloat16)
    k, v = torch.empty_like(q), torch.empty_like(q)
    L = _lib.lib()
    _lib.check(L.vog_tc_gemm_qkv_factored(_ptr(vis_lp), _rowmajor2d(vis_lp, 'vis_lp'), _ptr(wqkv_vis),
                                          _rowmajor2d(wqkv_vis, 'wqkv_vis'), M, K, tf32, n_heads, dhp, _ptr(lq),
                                          _rowmajor2d(lq, 'lq'), nfrm, nsrl, nppf2, _ptr(q), _ptr(k),
                                          _ptr(v), _stream()), 'vog_tc_gemm_qkv_factored')
    return q, k, v


def tc_gemm_gres(a, w, res_vis, res_lang, nfrm, nsrl, nppf2, bias=None, relu=False, lp_kind=LP_NONE,
                 want_f32=True, BN=None):
    """tc_gemm whose fp32 residual row for token m = (bt, s, p) is [res_vis[bt*nppf2+p] | res_lang[b*nsrl+s]]."""
    tf32 = _is_tf32(a, w)
    M, K = a.shape
    N = w.shape[0]
    _req(res_vis, torch.float32, 'res_vis', 2), _req(res_lang, torch.float32, 'res_lang', 2)
    dv = res_vis.shape[1]
    if w.shape[1] != K or dv + res_lang.shape[1] != N:
        raise ValueError('tc_gemm_gres: inconsistent shapes')
    out_f32 = torch.empty(M, N, device=a.device, dtype=torch.float32) if want_f32 else None
    out_lp = torch.empty(M, N, device=a.device, dtype=_LP_DTYPE[lp_kind]) if lp_kind != LP_NONE else None
    if bias is not None:
        _req(bias, torch.float32, 'bias', 1)
    BN = BN or pick_bn(N, M, _num_sms(a.device))
    if dv % BN:
        BN = pick_bn(N)
    L = _lib.lib()
    _lib.check(L.vog_tc_gemm_gres(_ptr(a), _rowmajor2d(a, 'a'), _ptr(w), _rowmajor2d(w, 'w'), M, N, K, tf32, BN,
                                  _ptr(bias), int(relu), _ptr(res_vis), _rowmajor2d(res_vis, 'res_vis'),
                                  _ptr(res_lang), _rowmajor2d(res_lang, 'res_lang'), dv, nfrm, nsrl, nppf2,
                                  _ptr(out_f32), N if out_f32 is not None else 0,
                                  _ptr(out_lp), N if out_lp is not None else 0, lp_kind, _stream()),
               'vog_tc_gemm_gres')
    return out_f32, out_lp


def tc_attn_fwd(q, k, v, N, head_dims, inv_scale, out_kind=LP_BF16, out=None, bias_mode=BIAS_NONE,
                a=None, nbox=0, bpe=None, dense=None, ak=None):
    """q,k,v [Bt,H,N,dhp] bf16 -> out [Bt*N, H*dhp] (bf16 or tf32-rounded fp32).  ak: the key factors of a rank-1 bias
    already expanded by pe_project_expand for exactly this (Bt, N, H, nbox, inv_scale) - skips the expansion pre-kernel."""
    for t, n in ((q, 'q'), (k, 'k'), (v, 'v')):
        _req(t, torch.bfloat16, n, 4)
        if not t.is_contiguous():
            raise ValueError(f'tc_attn_fwd: {n} must be contiguous')
    Bt, H, Nq, dhp = q.shape
    if Nq != N or k.shape != q.shape or v.shape != q.shape or len(head_dims) != H:
        raise ValueError('tc_attn_fwd: inconsistent shapes')
    if out is None:
        out = torch.empty(Bt * N, H * dhp, device=q.device, dtype=_LP_DTYPE[out_kind])
    dh_arr = (ctypes.c_int * H)(*head_dims)
    if bias_mode == BIAS_RANK1:
        _req(a, torch.float32, 'a', 2), _req(bpe, torch.float32, 'bpe', 1)
        if a.shape != (Bt * nbox, H) or not a.is_contiguous():
            raise ValueError(f'tc_attn_fwd: a must be contiguous [{Bt * nbox},{H}], got {tuple(a.shape)}')
    if bias_mode == BIAS_DENSE:
        _req(dense, torch.float32, 'dense', 4)
        if tuple(dense.shape) != (Bt, N, N, H) or not dense.is_contiguous():
            raise ValueError(f'tc_attn_fwd: dense bias must be contiguous [{Bt},{N},{N},{H}]')
    L = _lib.lib()
    ws, ws_bytes = None, 0
    if bias_mode == BIAS_RANK1:
        ws_bytes = L.vog_tc_attn_workspace_bytes(Bt, N, H)
        if ak is not None:
            if ak.dtype != torch.uint8 or ak.numel() < ws_bytes or ak.device != q.device:
                raise ValueError('tc_attn_fwd: ak does not belong to this attention call')
            ws, bias_mode = ak, BIAS_RANK1_EXPANDED
        else:
            ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
    _lib.check(L.vog_tc_attn_fwd(_ptr(q), _ptr(k), _ptr(v), Bt, N, H, dhp, dh_arr, float(inv_scale),
                                 bias_mode, _ptr(a), nbox, _ptr(bpe), _ptr(dense), _ptr(out),
                                 _rowmajor2d(out, 'out'), out_kind, _ptr(ws), ws_bytes, _stream()),
               'vog_tc_attn_fwd')
    return out


def lstm_layer_fwd(gx, whh, lens, T, Bq, kind, want_acts=False):
    """gx [T*Bq, 8H] fp32, whh [2,4H,H] fp32, lens [Bq] int64 -> h [T*Bq, 2H] (bf16 / tf32-rounded); with want_acts
    also the per-step activations [T*Bq, 2, 6, H] the backward consumes (vog_lstm_layer_fwd_train)."""
    _req(gx, torch.float32, 'gx', 2), _req(whh, torch.float32, 'whh', 3), _req(lens, torch.int64, 'lens', 1)
    H = whh.shape[2]
    if whh.shape != (2, 4 * H, H) or not whh.is_contiguous() or gx.shape != (T * Bq, 8 * H):
        raise ValueError('lstm_layer_fwd: inconsistent shapes')
    L = _lib.lib()
    out = torch.empty(T * Bq, 2 * H, device=gx.device, dtype=_LP_DTYPE.get(kind, torch.float32))
    ws = torch.empty(L.vog_lstm_workspace_bytes(Bq, H), device=gx.device, dtype=torch.uint8)
    if want_acts:
        acts = torch.empty(T * Bq, 2, 6, H, device=gx.device, dtype=torch.float32)
        _lib.check(L.vog_lstm_layer_fwd_train(_ptr(gx), _rowmajor2d(gx, 'gx'), _ptr(whh), _ptr(lens), T, Bq, H,
                                              _ptr(out), _rowmajor2d(out, 'out'), kind, _ptr(ws), _ptr(acts), _stream()),
                   'vog_lstm_layer_fwd_train')
        return out, acts
    _lib.check(L.vog_lstm_layer_fwd(_ptr(gx), _rowmajor2d(gx, 'gx'), _ptr(whh), _ptr(lens), T, Bq, H,
                                    _ptr(out), _rowmajor2d(out, 'out'), kind, _ptr(ws), _stream()),
               'vog_lstm_layer_fwd')
    return out


def lang_embed(words, mask, emb, pad_idx, kind):
    """words [Bq,nwords] int64, mask [Bq,T] int64 (-1 = pad), emb [V+1,E] fp32 -> x_lp [T*Bq, E] time-major."""
    _req(words, torch.int64, 'words', 2), _req(mask, torch.int64, 'mask', 2), _req(emb, torch.float32, 'emb', 2)
    words, mask, emb = words.contiguous(), mask.contiguous(), emb.contiguous()
    Bq, T = mask.shape
    out = torch.empty(T * Bq, emb.shape[1], device=emb.device, dtype=_LP_DTYPE.get(kind, torch.float32))
    L = _lib.lib()
    _lib.check(L.vog_lang_embed(_ptr(words), words.shape[1], _ptr(mask), T, _ptr(emb), emb.shape[1], int(pad_idx),
                                Bq, _ptr(out), kind, _stream()), 'vog_lang_embed')
    return out


def lang_gather(full, cap, T, Bq, kind):
    """full [T*Bq, D] fp32 time-major, cap [Bq,nsrl,2] int64 -> [Bq*nsrl, 2D] low precision."""
    _req(full, torch.float32, 'full', 2), _req(cap, torch.int64, 'cap', 3)
    full, cap = full.contiguous(), cap.contiguous()
    D, nsrl = full.shape[1], cap.shape[1]
    out = torch.empty(Bq * nsrl, 2 * D, device=full.device, dtype=_LP_DTYPE.get(kind, torch.float32))
    L = _lib.lib()
    _lib.check(L.vog_lang_gather(_ptr(full), D, _ptr(cap), T, Bq, nsrl, _ptr(out), kind, _stream()),
               'vog_lang_gather')
    return out


def mask_rows(x, msk, kind=LP_NONE):
    """x [rows, D] fp32, msk [rows] int64 -> x * msk (fp32) and, if kind is given, its low-precision copy."""
    _req(x, torch.float32, 'x', 2), _req(msk, torch.int64, 'msk', 1)
    x, msk = x.contiguous(), msk.contiguous()
    out = torch.empty_like(x)
    out_lp = torch.empty(x.shape, device=x.device, dtype=_LP_DTYPE[kind]) if kind != LP_NONE else None
    L = _lib.lib()
    _lib.check(L.vog_mask_rows(_ptr(x), _ptr(msk), x.shape[0], x.shape[1], _ptr(out), _ptr(out_lp), kind,
                               _stream()), 'vog_mask_rows')
    return out, out_lp


def build_xmul(vis, lang, B, nfrm, nsrl, nppf2, kind):
    """vis [B*nfrm*nppf2, dv] f32, lang [B*nsrl, dl] f32 -> x_mul [B*nfrm*nsrl*nppf2, dv+dl] f32 + lp copy."""
    _req(vis, torch.float32, 'vis', 2), _req(lang, torch.float32, 'lang', 2)
    if not (vis.is_contiguous() and lang.is_contiguous()):
        raise ValueError('build_xmul: contiguous inputs required')
    dv, dl = vis.shape[1], lang.shape[1]
    M = B * nfrm * nsrl * nppf2
    out = torch.empty(M, dv + dl, device=vis.device, dtype=torch.float32)
    out_lp = torch.empty(M, dv + dl, device=vis.device, dtype=_LP_DTYPE[kind])
    L = _lib.lib()
    _lib.check(L.vog_build_xmul(_ptr(vis), _ptr(lang), _ptr(out), _ptr(out_lp), kind, B, nfrm, nsrl, nppf2,
                                dv, dl, _stream()), 'vog_build_xmul')
    return out, out_lp


def tc_gemm_lin2(a, w, bias, w2, b2, srl_msk, cmp_msk, B, nfrm, nsrl, nppf2, ncmp, nppf, nfrm0, spat):
    """Scorer with its tail fused into the GEMM epilogue (see vog_tc_gemm_lin2) -> logits, scores [B,1,nsrl,P]."""
    tf32 = _is_tf32(a, w)
    M, K = a.shape
    N = w.shape[0]
    _req(bias, torch.float32, 'bias', 1), _req(srl_msk, torch.int64, 'srl_msk'), _req(cmp_msk, torch.int64, 'cmp_msk')
    if w.shape[1] != K or M != B * nfrm * nsrl * nppf2 or w2.numel() != N:
        raise ValueError('tc_gemm_lin2: inconsistent shapes')
    P = nfrm * nppf2
    logits = torch.empty(B, 1, nsrl, P, device=a.device, dtype=torch.float32)
    scores = torch.empty_like(logits)
    L = _lib.lib()
    _lib.check(L.vog_tc_gemm_lin2(_ptr(a), _rowmajor2d(a, 'a'), _ptr(w), _rowmajor2d(w, 'w'), M, N, K, tf32,
                                  _ptr(bias.contiguous()), _ptr(w2.contiguous()), _ptr(b2), _ptr(srl_msk.contiguous()),
                                  _ptr(cmp_msk.contiguous()), _ptr(logits), _ptr(scores), B, nfrm, nsrl, nppf2,
                                  ncmp, nppf, nfrm0, int(spat), _stream()), 'vog_tc_gemm_lin2')
    return logits, scores


def lin2_tail(h, w2, b2, srl_msk, cmp_msk, B, nfrm, nsrl, nppf2, ncmp, nppf, nfrm0, spat):
    """h [M,K] f32 -> logits, scores [B,1,nsrl,P] (see vog_lin2_tail)."""
    _req(h, torch.float32, 'h', 2), _req(srl_msk, torch.int64, 'srl_msk'), _req(cmp_msk, torch.int64, 'cmp_msk')
    P = nfrm * nppf2
    logits = torch.empty(B, 1, nsrl, P, device=h.device, dtype=torch.float32)
    scores = torch.empty_like(logits)
    L = _lib.lib()
    _lib.check(L.vog_lin2_tail(_ptr(h), _rowmajor2d(h, 'h'), _ptr(w2.contiguous()), _ptr(b2), _ptr(srl_msk.contiguous()),
                               _ptr(cmp_msk.contiguous()), _ptr(logits), _ptr(scores), B, nfrm, nsrl, nppf2,
                               h.shape[1], ncmp, nppf, nfrm0, int(spat), _stream()), 'vog_lin2_tail')
    return logits, scores


def loss_fwd(logits, props, gt, frm_mask, pnt_mask, srl_boxes, srl_lens, arg_boxes_mask, cmp_msk, target_cmp,
             ncmp, nppf, spat, loss_lambda=1.0, want_targets=False, keep_ws=False):
    """Grounding loss forward (see vog_loss_fwd): logits [B,nsrl,P] -> loss [1] f32 (and bool targets [B,nsrl,P])."""
    _req(logits, torch.float32, 'logits', 3), _req(props, torch.float32, 'props', 3), _req(gt, torch.float32, 'gt', 3)
    _req(frm_mask, torch.uint8, 'frm_mask', 3), _req(pnt_mask, torch.uint8, 'pnt_mask', 2)
    for t, n in ((srl_boxes, 'srl_boxes'), (srl_lens, 'srl_lens'), (arg_boxes_mask, 'arg_boxes_mask'),
                 (cmp_msk, 'cmp_msk'), (target_cmp, 'target_cmp')):
        _req(t, torch.int64, n)
    B, nsrl, P = logits.shape
    K, nb = gt.shape[1], srl_boxes.shape[-1]
    if props.shape[:2] != (B, P) or gt.shape[2] < 4 or frm_mask.shape != (B, P, K) or pnt_mask.shape != (B, P) or \
            srl_boxes.numel() != B * nsrl * nb or srl_lens.numel() != B * nsrl * nb:
        raise ValueError('loss_fwd: inconsistent shapes')
    logits, props, gt = logits.contiguous(), props.contiguous(), gt[:, :, :5].contiguous()
    L = _lib.lib()
    ws = torch.empty(L.vog_loss_workspace_bytes(B, nsrl, P), device=logits.device, dtype=torch.uint8)
    loss = torch.empty(1, device=logits.device, dtype=torch.float32)
    tg = torch.empty(B, nsrl, P, device=logits.device, dtype=torch.uint8) if want_targets else None
    _lib.check(L.vog_loss_fwd(_ptr(logits), _ptr(props), props.shape[2], _ptr(gt), _ptr(frm_mask.contiguous()),
                              _ptr(pnt_mask.contiguous()), _ptr(srl_boxes.contiguous()), _ptr(srl_lens.contiguous()),
                              _ptr(arg_boxes_mask.contiguous()), _ptr(cmp_msk.contiguous()),
                              _ptr(target_cmp.contiguous()), B, nsrl, nb, P, K, ncmp, nppf, int(spat),
                              float(loss_lambda), _ptr(tg), _ptr(ws), _ptr(loss), _stream()), 'vog_loss_fwd')
    if keep_ws:                      # for loss_bwd: the raw u8 targets and the workspace with mask + statistics
        return loss, tg, ws
    return (loss, tg.bool()) if want_targets else loss


def loss_bwd(logits, targets_u8, ws, grad_out):
    """d loss / d logits from the state a ``loss_fwd(..., want_targets=True, keep_ws=True)`` call left (see vog_loss_bwd)."""
    _req(logits, torch.float32, 'logits', 3), _req(targets_u8, torch.uint8, 'targets', 3), _req(grad_out, torch.float32, 'grad_out')
    B, nsrl, P = logits.shape
    grad = torch.empty_like(logits, memory_format=torch.contiguous_format)
    _lib.check(_lib.lib().vog_loss_bwd(_ptr(logits.contiguous()), _ptr(targets_u8), _ptr(ws), _ptr(grad_out.reshape(1).contiguous()),
                                       _ptr(grad), B, nsrl, P, _stream()), 'vog_loss_bwd')
    return grad
