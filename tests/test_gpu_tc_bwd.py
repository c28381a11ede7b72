"""Tensor-core kernels of the training step against float64 torch references on the same (bf16-rounded) operands:
vog_tc_gemm_tn (weight gradients), vog_tc_attn_fwd_train (log-sum-exp, dropout) and vog_tc_attn_bwd (dQ/dK/dV and the
rank-1 bias gradients).  The reference arithmetic is the autograd of the formula of code/transformer_code.py:136-160 with
the bias of code/mdl_vog.py:477-488; tolerances are set by the bf16 rounding of P / dS (documented per test)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'


def _ops():
    from vognet_pytorch_b200 import ops, ops_bwd
    return ops, ops_bwd


def _relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('K,N1,N2', [(64, 128, 64), (200, 128, 256), (1000, 256, 512), (333, 96, 192), (4100, 768, 768),
                                     (80, 8192, 1024), (16000, 256, 2048)])
def test_tc_gemm_tn(K, N1, N2):
    ops, ob = _ops()
    g = torch.Generator().manual_seed(K + N1)
    a = torch.randn(K, N1, generator=g).to(DEV).bfloat16()
    b = torch.randn(K, N2, generator=g).to(DEV).bfloat16()
    ref = a.double().t() @ b.double()
    out = ob.tc_gemm_tn(a, b)
    assert _relmax(out, ref) < 2e-5                    # fp32 accumulation of exact bf16 products
    out2 = ob.tc_gemm_tn(a, b, out=out)                # accumulates into a running gradient
    assert _relmax(out2, 2 * ref) < 2e-5
    # column-slice views (leading dimension > width)
    wide = torch.randn(K, N2 + 64, generator=g).to(DEV).bfloat16()
    out3 = ob.tc_gemm_tn(a, wide[:, 64:])
    assert _relmax(out3, a.double().t() @ wide[:, 64:].double()) < 2e-5


def _attn_case(Bt, N, nbox, head_dims, dhp, seed, rel=True):
    H = len(head_dims)
    g = torch.Generator().manual_seed(seed)
    def mk():
        t = torch.randn(Bt, H, N, dhp, generator=g) * 0.7
        for h, dh in enumerate(head_dims):
            t[:, h, :, dh:] = 0
        return t.to(DEV).bfloat16()
    q, k, v = mk(), mk(), mk()
    a = (torch.randn(Bt * nbox, H, generator=g) * 0.8).to(DEV) if rel else None
    bpe = (torch.randn(H, generator=g) * 0.3).to(DEV) if rel else None
    dout = torch.randn(Bt * N, H * dhp, generator=g)
    for h, dh in enumerate(head_dims):
        dout[:, h * dhp + dh:(h + 1) * dhp] = 0
    dout = dout.to(DEV).bfloat16()
    return q, k, v, a, bpe, dout


def _attn_ref(q, k, v, a, bpe, nbox, d_model, dout=None):
    """float64 autograd reference.  -> O [Bt*N, H*dhp], lse2 [Bt,H,N], and gradients if dout is given."""
    Bt, H, N, dhp = q.shape
    qd, kd, vd = (t.double().detach().requires_grad_(True) for t in (q, k, v))
    s = qd @ kd.transpose(-1, -2)
    ad = bd = None
    if a is not None:
        ad = a.double().detach().requires_grad_(True)
        bd = bpe.double().detach().requires_grad_(True)
        idx = torch.arange(N, device=q.device) % nbox
        at = ad.view(Bt, nbox, H)[:, idx].permute(0, 2, 1)                  # [Bt,H,N]
        s = s + torch.relu(at.unsqueeze(-1) - at.unsqueeze(-2) + bd.view(1, H, 1, 1))
    s = s / math.sqrt(d_model)
    lse2 = torch.logsumexp(s, -1) / math.log(2.0)
    o = (torch.softmax(s, -1) @ vd).permute(0, 2, 1, 3).reshape(Bt * N, H * dhp)
    if dout is None:
        return o.detach(), lse2.detach()
    o.backward(dout.double())
    return o.detach(), lse2.detach(), qd.grad, kd.grad, vd.grad, (ad.grad if ad is not None else None), \
        (bd.grad if bd is not None else None)


CASES = [
    # Bt, N, nbox, head_dims, dhp, rel
    (2, 100, 20, (171, 171, 170), 192, True),          # one tile, ragged, mul-like bias tiling (nsrl = 5)
    (1, 300, 300, (171, 171, 170), 192, True),         # three tiles, obj-like
    (3, 256, 64, (256, 256, 256), 256, True),          # exact tiles, dhp = 256
    (2, 200, 200, (128, 128), 128, False),             # plain Transformer (use_rel = False)
    (1, 640, 128, (256, 256, 256), 256, True),         # five key tiles: accumulator double buffering wraps
]


@pytest.mark.parametrize('Bt,N,nbox,head_dims,dhp,rel', CASES)
def test_tc_attn_train_forward_lse(Bt, N, nbox, head_dims, dhp, rel):
    ops, ob = _ops()
    d_model = sum(head_dims)
    q, k, v, a, bpe, _ = _attn_case(Bt, N, nbox, head_dims, dhp, 1, rel)
    o_ref, lse_ref = _attn_ref(q, k, v, a, bpe, nbox, d_model)
    kw = dict(bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe) if rel else {}
    out, lse = ob.tc_attn_fwd_train(q, k, v, N, head_dims, 1.0 / math.sqrt(d_model), **kw)
    assert float((lse.double() - lse_ref).abs().max()) < 2e-3             # ex2.approx + bf16 P do not enter the lse
    assert _relmax(out, o_ref) < 1.5e-2                                    # bf16 P and bf16 output
    # same entry point as the inference kernel
    out_inf = ops.tc_attn_fwd(q, k, v, N, head_dims, 1.0 / math.sqrt(d_model), **kw)
    assert torch.equal(out_inf, out)


@pytest.mark.parametrize('Bt,N,nbox,head_dims,dhp,rel', CASES)
def test_tc_attn_bwd(Bt, N, nbox, head_dims, dhp, rel):
    ops, ob = _ops()
    H, d_model = len(head_dims), sum(head_dims)
    inv = 1.0 / math.sqrt(d_model)
    q, k, v, a, bpe, dout = _attn_case(Bt, N, nbox, head_dims, dhp, 2, rel)
    _, _, dq, dk, dv, da_ref, db_ref = _attn_ref(q, k, v, a, bpe, nbox, d_model, dout)
    kw = dict(bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe) if rel else {}
    out, lse = ob.tc_attn_fwd_train(q, k, v, N, head_dims, inv, **kw)
    da = torch.zeros_like(a) if rel else None
    dbpe = torch.zeros_like(bpe) if rel else None
    dqkv = ob.tc_attn_bwd(q, k, v, out, dout, lse, N, head_dims, inv, da=da, dbpe=dbpe, **kw)
    got = dqkv.view(Bt, N, 3, H, dhp).permute(2, 0, 3, 1, 4)              # [3,Bt,H,N,dhp]
    # P and dS are rounded to bf16 before the three contractions and the results are stored as bf16: 2e-2 of the
    # largest entry (measured 4-8e-3)
    for name, g_, r_ in (('dq', got[0], dq), ('dk', got[1], dk), ('dv', got[2], dv)):
        assert _relmax(g_, r_) < 2e-2, name
    if rel:
        assert _relmax(da, da_ref) < 5e-3                                  # reduced from the fp32 values, not bf16
        assert _relmax(dbpe, db_ref) < 5e-3
        # accumulation semantics
        dqkv2 = ob.tc_attn_bwd(q, k, v, out, dout, lse, N, head_dims, inv, da=da, dbpe=dbpe, **kw)
        assert torch.equal(dqkv2, dqkv)
        assert _relmax(da, 2 * da_ref) < 5e-3


def test_tc_attn_dropout_is_consistent_and_calibrated():
    """Dropout on the probabilities: (i) keep rate = 1 - p and E[out] = out without dropout; (ii) the backward
    regenerates the forward's mask: gradients equal the float64 gradients of the SAME masked function, with the mask
    recovered from the kernel itself (V = identity makes the forward output the dropped probability matrix)."""
    ops, ob = _ops()
    Bt, N, H, dhp, p = 1, 128, 2, 128, 0.25
    head_dims = (128, 128)
    d_model = 256
    inv = 1.0 / math.sqrt(d_model)
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(Bt, H, N, dhp, generator=g) * 0.5).to(DEV).bfloat16()
    k = (torch.randn(Bt, H, N, dhp, generator=g) * 0.5).to(DEV).bfloat16()
    eye = torch.eye(N, dhp).expand(Bt, H, N, dhp).contiguous().to(DEV).bfloat16()
    out0, lse0 = ob.tc_attn_fwd_train(q, k, eye, N, head_dims, inv)
    out1, lse1 = ob.tc_attn_fwd_train(q, k, eye, N, head_dims, inv, drop_p=p, seed=1234)
    out2, _ = ob.tc_attn_fwd_train(q, k, eye, N, head_dims, inv, drop_p=p, seed=1234)
    out3, _ = ob.tc_attn_fwd_train(q, k, eye, N, head_dims, inv, drop_p=p, seed=99)
    assert torch.equal(out1, out2) and not torch.equal(out1, out3)          # a pure function of the seed
    assert torch.equal(lse0, lse1)                                          # the row sum keeps every probability
    P0 = out0.view(Bt, N, H, dhp).permute(0, 2, 1, 3).float()               # [Bt,H,N,N] probabilities
    P1 = out1.view(Bt, N, H, dhp).permute(0, 2, 1, 3).float()
    keep = P1 != 0
    big = P0 > 1e-3                                                         # entries whose bf16 value cannot round to 0
    rate = float(keep[big].float().mean())
    n = int(big.sum())
    assert abs(rate - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n), (rate, n)
    assert float((P1[keep & big] / P0[keep & big] - 1 / (1 - p)).abs().max()) < 2e-2
    # rows and columns are not correlated: keep rate per row / per column within binomial bounds
    for dim in (-1, -2):
        r = keep.float().mean(dim)
        assert float((r - (1 - p)).abs().max()) < 6 * math.sqrt(p * (1 - p) / N)
    # ---- backward with the recovered mask
    v = (torch.randn(Bt, H, N, dhp, generator=g) * 0.5).to(DEV).bfloat16()
    dout = torch.randn(Bt * N, H * dhp, generator=g).to(DEV).bfloat16()
    out, lse = ob.tc_attn_fwd_train(q, k, v, N, head_dims, inv, drop_p=p, seed=1234)
    dqkv = ob.tc_attn_bwd(q, k, v, out, dout, lse, N, head_dims, inv, drop_p=p, seed=1234)
    qd, kd, vd = (t.double().detach().requires_grad_(True) for t in (q, k, v))
    P = torch.softmax(qd @ kd.transpose(-1, -2) * inv, -1) * keep.double() / (1 - p)
    o = (P @ vd).permute(0, 2, 1, 3).reshape(Bt * N, H * dhp)
    assert _relmax(out, o.detach()) < 1.5e-2
    o.backward(dout.double())
    got = dqkv.view(Bt, N, 3, H, dhp).permute(2, 0, 3, 1, 4)
    for name, g_, r_ in (('dq', got[0], qd.grad), ('dk', got[1], kd.grad), ('dv', got[2], vd.grad)):
        assert _relmax(g_, r_) < 2e-2, name


def test_fp32_and_tcgen05_attention_draw_the_same_dropout_mask():
    """vog_attn_fwd_f32 / vog_attn_bwd_f32 with drop_p use the same counter-based mask as the tensor-core kernels:
    same seed -> same dropped probabilities (outputs agree to bf16 rounding), forward and backward."""
    ops, ob = _ops()
    Bt, N, nbox, head_dims, dhp = 2, 150, 30, (171, 171, 170), 192
    H, d = len(head_dims), sum(head_dims)
    inv = 1.0 / math.sqrt(d)
    q, k, v, a, bpe, dout = _attn_case(Bt, N, nbox, head_dims, dhp, 7, True)
    kw = dict(bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe)
    out_tc, lse_tc = ob.tc_attn_fwd_train(q, k, v, N, head_dims, inv, drop_p=0.2, seed=99, **kw)
    da_tc, db_tc = torch.zeros_like(a), torch.zeros_like(bpe)
    dqkv_tc = ob.tc_attn_bwd(q, k, v, out_tc, dout, lse_tc, N, head_dims, inv, da=da_tc, dbpe=db_tc, drop_p=0.2, seed=99, **kw)

    def unpad(t):                                   # [Bt,H,N,dhp] -> [Bt*N, d]
        return torch.cat([t[:, h, :, :dh] for h, dh in enumerate(head_dims)], -1).reshape(Bt * N, d).float().contiguous()
    qf, kf, vf = unpad(q), unpad(k), unpad(v)
    dof = torch.cat([dout.view(Bt * N, H, dhp)[:, h, :dh] for h, dh in enumerate(head_dims)], -1).float().contiguous()
    lse = torch.empty(Bt * H * N, device=DEV)
    out_f = ops.attn_fwd_f32(qf, kf, vf, Bt, N, head_dims, inv, lse=lse, drop_p=0.2, seed=99, **kw)
    out_f0 = ops.attn_fwd_f32(qf, kf, vf, Bt, N, head_dims, inv)
    got = torch.cat([out_tc.view(Bt * N, H, dhp)[:, h, :dh] for h, dh in enumerate(head_dims)], -1).float()
    assert _relmax(got, out_f) < 1.5e-2                         # same mask
    assert _relmax(out_f0, out_f) > 1e-1                        # ... and the mask matters
    da_f, db_f = torch.zeros_like(a), torch.zeros_like(bpe)
    dqkv_f, _ = ob.attn_bwd_f32(qf, kf, vf, out_f, dof, lse, Bt, N, head_dims, inv, da=da_f, dbpe=db_f, drop_p=0.2, seed=99, **kw)
    tc = dqkv_tc.view(Bt * N, 3, H, dhp)
    for i, name in enumerate(('dq', 'dk', 'dv')):
        g_tc = torch.cat([tc[:, i, h, :dh] for h, dh in enumerate(head_dims)], -1).float()
        assert _relmax(g_tc, dqkv_f[:, i * d:(i + 1) * d]) < 2.5e-2, name
    assert _relmax(da_tc, da_f) < 1e-2 and _relmax(db_tc, db_f) < 1e-2
