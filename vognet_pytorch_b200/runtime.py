"""Host-side runtime helpers around the forward: query sharding across ranks (one process per GPU,
no data-path collective in the forward - SURVEY.md section 8e), pinned-host -> device batch
prefetch on a copy stream, and the NCCL/gloo gather of predictions that replaces the reference's
per-rank pickle files (code/eval_vsrl_corr.py:125-140)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) slice of n queries owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world):
    """Rank-local view of a batch dict: every tensor is sliced on its first (query) axis."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def gather_predictions(pred, world=None, group=None):
    """all_gather a dict of per-rank prediction tensors (equal trailing shapes, possibly different
    query counts) back into global query order.  Works with NCCL (CUDA tensors) and gloo (CPU)."""
    if not (dist.is_available() and dist.is_initialized()):
        return pred
    world = world or dist.get_world_size(group)
    out = {}
    for k, v in pred.items():
        n_local = torch.tensor([v.shape[0]], device=v.device, dtype=torch.int64)
        counts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(counts, n_local, group=group)
        counts = [int(c.item()) for c in counts]
        nmax = max(counts)
        pad = torch.zeros((nmax,) + tuple(v.shape[1:]), device=v.device, dtype=v.dtype)
        pad[:v.shape[0]] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
    return out


def concat_batch(sep_batch, conc_type, nfrm, nppf, sentence_slot=0):
    """SPAT / TEMP batch dict from a per-video (SEP-layout) batch that is already on the GPU: the visual tensors go
    through ``vog_concat_videos`` (what the reference's dataset does per sample on the host,
    code/dat_loader_simple.py:1067-1207,1231-1292), the language tensors keep the one sentence slot the concatenated
    models read (``[B,1,...]``; SPAT/TEMP samples carry the query sentence once, append_everywhere=False :934-941)."""
    from . import ops
    feat, seg, props = ops.concat_videos(sep_batch['pad_region_feature'], sep_batch['seg_feature_for_frms'],
                                         sep_batch['pad_proposals'], conc_type, nfrm, nppf)
    out = {'pad_region_feature': feat, 'seg_feature_for_frms': seg, 'pad_proposals': props,
           'new_srl_idxs': sep_batch['new_srl_idxs'], 'num_cmp_msk': sep_batch['num_cmp_msk']}
    for k in ('srl_arg_words_ind', 'srl_arg_word_mask', 'srl_tag_word_ind', 'srl_arg_word_mask_len',
              'srl_arg_words_capture', 'srl_arg_inds_msk'):
        if k in sep_batch:
            v = sep_batch[k]
            out[k] = v[:, sentence_slot:sentence_slot + 1].contiguous() if v.shape[1] > 1 else v
    return out


def allreduce_flat_sum_(flat, group=None, async_op=False):
    """In-place SUM of one flat gradient buffer over the ranks - a single collective for the whole model - and the
    factor that turns it into the mean DistributedDataParallel produces (code/main_dist.py:76-85).
    -> (1/world, work handle or None).  With ``async_op`` the collective is only enqueued (NCCL: on its own stream);
    ``work.wait()`` orders the caller's current stream after it."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0, None
    work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return 1.0 / dist.get_world_size(group), (work if async_op else None)


def max_over_ranks(seconds, device):
    """Device-side max of a per-rank duration (the multi-GPU timing rule of bench.py)."""
    t = torch.tensor([seconds], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class BatchPrefetcher:
    """Double-buffered host -> device staging: while the model works on batch i (compute stream),
    batch i+1 is copied from pinned host memory on a dedicated copy stream.  ``next()`` returns a
    device batch dict whose copies are ordered before the caller's current stream."""

    def __init__(self, host_batches, device, depth=2):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.queue = []
        for _ in range(depth):
            self._enqueue()

    def _enqueue(self):
        try:
            hb = next(self.it)
        except StopIteration:
            return
        with torch.cuda.stream(self.copy_stream):
            db = {k: v.to(self.device, non_blocking=True) for k, v in hb.items()}
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((db, ev))

    def next(self):
        if not self.queue:
            return None
        db, ev = self.queue.pop(0)
        torch.cuda.current_stream(self.device).wait_event(ev)
        for v in db.values():                       # the compute stream now owns these buffers
            v.record_stream(torch.cuda.current_stream(self.device))
        self._enqueue()
        return db
