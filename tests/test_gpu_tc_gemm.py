"""tcgen05 GEMM (vog_tc_gemm / vog_tc_gemm_qkv) against a float64 product of the SAME rounded
operands: the only differences left are fp32 accumulation order, so the tolerance is tight."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from vognet_pytorch_b200 import ops      # noqa: E402

DEV = 'cuda:0'


def _u(shape, seed, lo=-1.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * (hi - lo) + lo


def _tf32(x):
    i = x.view(torch.int32)
    return (((i + 0x1000) & ~0x1FFF)).view(torch.float32)


@pytest.mark.parametrize('M,N,K,BN', [(128, 128, 64, 128), (128, 256, 512, 256), (300, 256, 2048, 256),
                                      (1000, 768, 768, 256), (77, 384, 768, 128), (4000, 512, 576, 256),
                                      (200, 192, 64, 192), (129, 64, 8, 64), (5, 32, 3072, 32)])
@pytest.mark.parametrize('mode', ['bf16', 'tf32'])
def test_tc_gemm_plain(M, N, K, BN, mode):
    a, w = _u((M, K), 1), _u((N, K), 2)
    if mode == 'bf16':
        al, wl = a.bfloat16(), w.bfloat16()
    else:
        al, wl = _tf32(a), _tf32(w)
    ref = al.double() @ wl.double().t()
    out, _ = ops.tc_gemm(al.to(DEV), wl.to(DEV), BN=BN)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, (K / 64) ** 0.5), err


def test_tc_gemm_epilogue_all_options():
    M, N, K, rep = 150, 256, 320, 3
    a, w, b = _u((M, K), 3).bfloat16(), _u((N, K), 4).bfloat16(), _u((N,), 5)
    res = _u((M, N), 6)
    ref = torch.relu(a.double() @ w.double().t() + b.double()) + res.double()
    o32 = torch.zeros(M * rep, 300, device=DEV)
    olp = torch.zeros(M * rep, 264, device=DEV, dtype=torch.bfloat16)
    ops.tc_gemm(a.to(DEV), w.to(DEV), bias=b.to(DEV), residual=res.to(DEV), relu=True,
                out_f32=o32[:, 20:276], out_lp=olp[:, 8:264], rep=rep)
    got = o32[:, 20:276].cpu().view(M, rep, N)
    for r in range(rep):
        assert (got[:, r].double() - ref).abs().max() < 3e-4
    assert o32[:, :20].abs().max() == 0 and o32[:, 276:].abs().max() == 0
    lp = olp[:, 8:264].cpu().view(M, rep, N)
    assert ((lp[:, 1].double() - ref).abs() / (1 + ref.abs())).max() < 8e-3     # bf16 rounding of values up to ~16
    assert torch.equal(lp[:, 0], got[:, 0].bfloat16())
    # tf32-rounded low precision copy
    _, o2 = ops.tc_gemm(a.to(DEV), w.to(DEV), lp_kind=ops.LP_TF32, want_f32=False)
    assert torch.equal(o2.cpu(), _tf32((a.double() @ w.double().t()).float())) or \
        (o2.cpu().double() - a.double() @ w.double().t()).abs().max() < 2e-2


@pytest.mark.parametrize('Bt,N,d,H', [(2, 200, 512, 3), (3, 100, 768, 3), (1, 37, 512, 6), (2, 128, 512, 3)])
def test_tc_gemm_qkv_scatter(Bt, N, d, H):
    hd = ops.chunk_sizes(d, H)
    dhp = ops.round_up(max(hd), 64)
    x = _u((Bt * N, d), 7).bfloat16()
    ws = [_u((d, d), 8 + i) for i in range(3)]
    wp = torch.zeros(3 * H * dhp, d)
    for i, wm in enumerate(ws):
        off = 0
        for h, dh in enumerate(hd):
            wp[(i * H + h) * dhp:(i * H + h) * dhp + dh] = wm[off:off + dh]
            off += dh
    wp = wp.bfloat16()
    q, k, vt = ops.tc_gemm_qkv(x.to(DEV), wp.to(DEV), Bt, N, H, dhp)
    torch.cuda.synchronize()
    full = (x.double() @ wp.double().t()).view(Bt, N, 3, H, dhp)
    for got, which in ((q, 0), (k, 1)):
        ref = full[:, :, which].permute(0, 2, 1, 3)
        assert ((got.cpu().double() - ref).abs() / (1 + ref.abs())).max() < 8e-3      # bf16 outputs
    refv = full[:, :, 2].permute(0, 2, 1, 3)                     # [Bt,H,N,dhp]
    assert ((vt.cpu().double() - refv).abs() / (1 + refv.abs())).max() < 8e-3
    # padded head columns are exact zeros
    for h, dh in enumerate(hd):
        if dh < dhp:
            assert q[:, h, :, dh:].abs().max() == 0 and vt[:, h, :, dh:].abs().max() == 0


@pytest.mark.parametrize('B,nfrm,nsrl,nppf2,mode', [(2, 10, 5, 20, 'bf16'), (1, 3, 5, 7, 'tf32'), (2, 4, 3, 100, 'bf16'),
                                                      (4, 10, 5, 20, 'tf32')])
def test_factored_qkv_and_gathered_residual_match_materialised_tokens(B, nfrm, nsrl, nppf2, mode):
    """vog_tc_gemm_qkv_factored / vog_tc_gemm_gres (the [vis|lang] token matrix is never written)
    against the same GEMMs on the materialised token matrix (vog_build_xmul + vog_tc_gemm_qkv /
    vog_tc_gemm with a plain residual) and against float64."""
    dv, dl, H = 512, 256, 3
    d = dv + dl
    dhp = 256
    kind = ops.LP_BF16 if mode == 'bf16' else ops.LP_TF32
    Bt, N = B * nfrm, nsrl * nppf2
    vis = _u((Bt * nppf2, dv), 21).to(DEV)
    lang = _u((B * nsrl, dl), 22).to(DEV)
    wqkv = ops.cast_lp((_u((3 * H * dhp, d), 23) * 0.05).to(DEV), kind)
    vis_lp, lang_lp = ops.cast_lp(vis, kind), ops.cast_lp(lang, kind)
    xm, xm_lp = ops.build_xmul(vis, lang, B, nfrm, nsrl, nppf2, kind)
    q0, k0, vt0 = ops.tc_gemm_qkv(xm_lp, wqkv, Bt, N, H, dhp)
    lq, _ = ops.tc_gemm(lang_lp, wqkv[:, dv:])
    q1, k1, vt1 = ops.tc_gemm_qkv_factored(vis_lp, wqkv[:, :dv], lq, Bt, nfrm, nsrl, nppf2, H, dhp)
    torch.cuda.synchronize()
    full = (xm_lp.double().cpu() @ wqkv.double().cpu().t()).view(Bt, N, 3, H, dhp)
    for got0, got1, which in ((q0, q1, 0), (k0, k1, 1)):
        ref = full[:, :, which].permute(0, 2, 1, 3)
        assert ((got1.cpu().double() - ref).abs() / (1 + ref.abs())).max() < 8e-3
        assert ((got1.float() - got0.float()).abs() / (1 + got0.float().abs())).max() < 8e-3
    refv = full[:, :, 2].permute(0, 2, 1, 3)
    assert ((vt1.cpu().double() - refv).abs() / (1 + refv.abs())).max() < 8e-3
    assert ((vt1.float() - vt0.float()).abs() / (1 + vt0.float().abs())).max() < 8e-3
    # gathered residual
    a = ops.cast_lp(_u((Bt * N, H * dhp), 24).to(DEV), kind)
    wo = ops.cast_lp((_u((d, H * dhp), 25) * 0.05).to(DEV), kind)
    ref_out, _ = ops.tc_gemm(a, wo, residual=xm)
    got_out, _ = ops.tc_gemm_gres(a, wo, vis, lang, nfrm, nsrl, nppf2)
    torch.cuda.synchronize()
    assert torch.allclose(ref_out, got_out, atol=2e-4, rtol=0)      # split-K vs single pass: summation order only


@pytest.mark.parametrize('B,nfrm,nsrl,nppf2,ncmp,spat,mode', [(2, 10, 5, 20, 4, True, 'tf32'), (2, 40, 5, 5, 4, False, 'tf32'),
                                                                (1, 10, 5, 400, 4, True, 'bf16'), (3, 10, 2, 7, 1, True, 'bf16')])
def test_fused_scorer_tail_matches_gemm_plus_tail(B, nfrm, nsrl, nppf2, ncmp, spat, mode):
    """vog_tc_gemm_lin2 (lin2[2] + un-regroup + sigmoid*masks inside the GEMM epilogue) against
    vog_tc_gemm followed by vog_lin2_tail, and against float64."""
    kind = ops.LP_BF16 if mode == 'bf16' else ops.LP_TF32
    nfrm0 = 10
    nppf = nppf2 // ncmp if spat else nppf2
    M, K, N = B * nfrm * nsrl * nppf2, 768, 256
    a = ops.cast_lp(_u((M, K), 31).to(DEV), kind)
    w1 = ops.cast_lp((_u((N, K), 32) * 0.05).to(DEV), kind)
    b1, w2, b2 = _u((N,), 33).to(DEV), _u((1, N), 34).to(DEV), _u((1,), 35).to(DEV)
    g = torch.Generator().manual_seed(36)
    srl = torch.randint(0, 2, (B, nsrl), generator=g).to(DEV)
    cmp_ = torch.randint(0, 2, (B, ncmp), generator=g).to(DEV)
    h, _ = ops.tc_gemm(a, w1, bias=b1, relu=True)
    lg0, sc0 = ops.lin2_tail(h, w2, b2, srl, cmp_, B, nfrm, nsrl, nppf2, ncmp, nppf, nfrm0, spat)
    lg1, sc1 = ops.tc_gemm_lin2(a, w1, b1, w2, b2, srl, cmp_, B, nfrm, nsrl, nppf2, ncmp, nppf, nfrm0, spat)
    torch.cuda.synchronize()
    assert torch.allclose(lg0, lg1, atol=2e-4, rtol=0)
    assert torch.allclose(sc0, sc1, atol=1e-4, rtol=0)
    href = torch.relu(a.double().cpu() @ w1.double().cpu().t() + b1.double().cpu())
    lref = (href @ w2.double().cpu().t()).view(B, nfrm, nsrl, nppf2).permute(0, 2, 1, 3).reshape(B, 1, nsrl, nfrm * nppf2) + b2.double().cpu()
    assert (lg1.cpu().double() - lref).abs().max() < 1e-3
