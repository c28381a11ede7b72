"""world_size-2 gloo test of the host-side multi-GPU logic: query sharding (no data-path collective
in the forward), gather of per-rank predictions back into global order, max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vognet_pytorch_b200 import runtime, synth


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nq, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batch = synth.make_batch('spat', B=nq, ncmp=4, nppf=5, seed=7)      # identical on every rank
        mine = runtime.shard_batch(batch, rank, world)
        lo, hi = runtime.shard_range(nq, rank, world)
        assert mine['pad_proposals'].shape[0] == hi - lo
        assert torch.equal(mine['pad_proposals'], batch['pad_proposals'][lo:hi])
        # stand-in for the per-rank forward + selection: a deterministic function of the shard
        pred = {'scores': mine['pad_proposals'][:, :50, 6].clone(),
                'indexs': mine['srl_arg_word_mask_len'].clone()}
        full = runtime.gather_predictions(pred)
        assert torch.equal(full['scores'], batch['pad_proposals'][:, :50, 6])
        assert torch.equal(full['indexs'], batch['srl_arg_word_mask_len'])
        # flat gradient all-reduce of the data-parallel training step: one collective, mean = sum * 1/world
        g = torch.arange(10, dtype=torch.float32) * (rank + 1)
        scale, _ = runtime.allreduce_flat_sum_(g)
        assert scale == 1.0 / world and torch.equal(g * scale, torch.arange(10, dtype=torch.float32) * 1.5)
        t = runtime.max_over_ranks(1.0 + rank, 'cpu')
        assert t == float(world)
        ret[rank] = 'ok'
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('nq', [8, 5])          # even and ragged split
def test_shard_gather_world2(nq):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [mp.get_context('spawn').Process(target=_worker, args=(r, world, port, nq, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(ret) == {0: 'ok', 1: 'ok'}


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 33):
        for w in (1, 2, 3, 8):
            spans = [runtime.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
