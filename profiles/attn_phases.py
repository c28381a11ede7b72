#!/usr/bin/env python
"""In-kernel clock64 phase breakdown of one softmax warp of the attention kernel (debug hook)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vognet_pytorch_b200 import ops, _lib
dev = 'cuda:0'
L = _lib.lib()
L.vog_debug_attn_prof.argtypes = [ctypes.c_void_p]
L.vog_debug_attn_prof.restype = None
names7 = 'rescale chk + next bias/score loads (waits S_{j+1})'
names = ['wait tmem ld (prefetched)', 'reg copy + next bias loads', 'scale/bias/max', 'pair exchange', 'rescale chk + exp + sum', 'wait P buffer', 'pack+sts+fence+arrive']
for Bt, N, d in ((40, 2000, 768), (4, 4000, 512)):
    hd = ops.chunk_sizes(d, 3); dhp = ops.round_up(max(hd), 64)
    q = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    k = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    vt = (torch.rand(Bt, 3, N, dhp, device=dev) - 0.5).bfloat16()
    nbox = N // 5
    a = torch.rand(Bt * nbox, 3, device=dev); bpe = torch.zeros(3, device=dev)
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    for _ in range(2):
        ops.tc_attn_fwd(q, k, vt, N, hd, 1.0 / d ** 0.5, bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe)
    L.vog_debug_attn_prof(ctypes.c_void_p(buf.data_ptr()))
    ops.tc_attn_fwd(q, k, vt, N, hd, 1.0 / d ** 0.5, bias_mode=ops.BIAS_RANK1, a=a, nbox=nbox, bpe=bpe)
    torch.cuda.synchronize()
    L.vog_debug_attn_prof(None)
    v = buf.cpu().tolist(); T = v[12]; v[4] += 0
    print(f'Bt={Bt} N={N} d={d}: {T} tiles, {sum(v[:8]) / T:.0f} cycles/tile')
    for n, c in zip(names, v[:7]):
        print(f'   {n:28s} {c / T:8.0f}')
    print(f'   {names7:28s} {v[7] / T:8.0f}   (split out of the exp phase above)')
    print(f'  MMA thread: {(sum(v[8:12]) + v[13]) / T:.0f} cycles/tile')
    for n, c in zip(['wait K (k_full)', 'issue S MMAs + commits', 'wait V + P (p_full)', 'issue PV MMAs + commits'], v[8:12]):
        print(f'   {n:28s} {c / T:8.0f}')
    print(f'   of which wait V (v_full)     {v[13] / T:8.0f}   (the rest of "wait V + P" is the wait for P_j)')
    if v[32]:
        t0 = v[32]
        print('   tile: K requested / MMA starts waiting / K observed (cycles since the first K request)')
        for j in range(min(int(T), 14)):
            print(f'     {j:3d}: {v[32 + j] - t0:8d} {v[64 + j] - t0:8d} {v[96 + j] - t0:8d}   request->observed {v[96 + j] - v[32 + j]:6d}')
