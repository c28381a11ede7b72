#!/usr/bin/env python
"""Run ONE hot-path op a few times (for `ncu --set full -k regex:<kernel>` captures).
usage: one_op.py gemm M N K [bf16|tf32] [residual] | attn Bt N d | attn_bwd Bt N d [drop] | lstm"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops  # noqa: E402

dev = 'cuda:0'
what = sys.argv[1]
if what == 'gemm':
    M, N, K = [int(v) for v in sys.argv[2:5]]
    kind = ops.LP_TF32 if 'tf32' in sys.argv else ops.LP_BF16
    dt = torch.bfloat16 if kind == ops.LP_BF16 else torch.float32
    a = (torch.rand(M, K, device=dev) - 0.5).to(dt)
    w = (torch.rand(N, K, device=dev) - 0.5).to(dt)
    r = torch.rand(M, N, device=dev) if 'residual' in sys.argv else None
    o32 = torch.empty(M, N, device=dev) if r is not None else None
    olp = torch.empty(M, N, device=dev, dtype=dt) if r is None else None
    for _ in range(4):
        ops.tc_gemm(a, w, residual=r, out_f32=o32, out_lp=olp, want_f32=r is not None,
                    lp_kind=(kind if r is None else ops.LP_NONE))
elif what == 'attn':
    Bt, N, d = [int(v) for v in sys.argv[2:5]]
    hd = ops.chunk_sizes(d, 3)
    dhp = ops.round_up(max(hd), 64)
    q = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    k = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    vt = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    nbox = N // 5 if N % 5 == 0 else N
    a = torch.rand(Bt * nbox, 3, device=dev)
    bpe = torch.zeros(3, device=dev)
    for _ in range(4):
        ops.tc_attn_fwd(q, k, vt, N, hd, 1.0 / d ** 0.5, bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe)
elif what == 'attn_bwd':                   # training forward + backward of the fused attention (spat/p100: 40 2000 768, 4 4000 512)
    from vognet_pytorch_b200 import ops_bwd as ob
    Bt, N, d = [int(v) for v in sys.argv[2:5]]
    drop = 0.2 if 'drop' in sys.argv else 0.0
    hd = ops.chunk_sizes(d, 3)
    dhp = ops.round_up(max(hd), 64)
    q = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    k = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    vt = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    do = (torch.rand(Bt * N, 3 * dhp, device=dev) - 0.5).bfloat16()
    nbox = N // 5 if N % 5 == 0 else N
    a = torch.rand(Bt * nbox, 3, device=dev)
    bpe = torch.zeros(3, device=dev)
    da, dbpe = torch.zeros_like(a), torch.zeros_like(bpe)
    kw = dict(bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe, drop_p=drop, seed=7)
    ts = []
    for it in range(4):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out, lse = ob.tc_attn_fwd_train(q, k, vt, N, hd, 1.0 / d ** 0.5, **kw)
        e[1].record()
        ob.tc_attn_bwd(q, k, vt, out, do, lse, N, hd, 1.0 / d ** 0.5, da=da, dbpe=dbpe, **kw)
        e[2].record()
        torch.cuda.synchronize()
        ts.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
    fl = 4.0 * Bt * N * N * d
    print(f'attn_bwd Bt={Bt} N={N} d={d} drop={drop}: fwd {ts[-1][0]:.3f} ms ({fl / ts[-1][0] / 1e9:.0f} TF/s), '
          f'bwd {ts[-1][1]:.3f} ms ({2.5 * fl / ts[-1][1] / 1e9:.0f} TF/s)')
elif what == 'qkvf':                       # factorised mul_tx QKV at spat/p100: B=4, nfrm=10, nsrl=5, nppf2=400
    B, nfrm, nsrl, nppf2 = [int(v) for v in sys.argv[2:6]]
    kind = ops.LP_TF32 if 'tf32' in sys.argv else ops.LP_BF16
    Bt = B * nfrm
    vis = ops.cast_lp(torch.rand(Bt * nppf2, 512, device=dev) - 0.5, kind)
    lq = torch.rand(B * nsrl, 2304, device=dev) - 0.5
    w = ops.cast_lp((torch.rand(2304, 768, device=dev) - 0.5) * 0.05, kind)
    for _ in range(4):
        ops.tc_gemm_qkv_factored(vis, w[:, :512], lq, Bt, nfrm, nsrl, nppf2, 3, 256)
elif what == 'gres':                       # mul_tx wo GEMM with the gathered [vis|lang] residual
    B, nfrm, nsrl, nppf2 = [int(v) for v in sys.argv[2:6]]
    kind = ops.LP_TF32 if 'tf32' in sys.argv else ops.LP_BF16
    Bt = B * nfrm
    M = Bt * nsrl * nppf2
    a = ops.cast_lp(torch.rand(M, 768, device=dev) - 0.5, kind)
    w = ops.cast_lp((torch.rand(768, 768, device=dev) - 0.5) * 0.05, kind)
    vis = torch.rand(Bt * nppf2, 512, device=dev)
    lang = torch.rand(B * nsrl, 256, device=dev)
    for _ in range(4):
        ops.tc_gemm_gres(a, w, vis, lang, nfrm, nsrl, nppf2)
elif what == 'lstm':                       # one recurrence layer: T=20, Bq=4, H=1024, lengths as in the gt5 workload
    T, Bq, H = 20, 4, 1024
    kind = ops.LP_TF32
    gx = torch.rand(T * Bq, 8 * H, device=dev) - 0.5
    whh = (torch.rand(2, 4 * H, H, device=dev) - 0.5) / 32
    lens = torch.tensor([20, 20, 20, 20], device=dev)
    for _ in range(4):
        ops.lstm_layer_fwd(gx, whh, lens, T, Bq, kind)
torch.cuda.synchronize()
print('done')
