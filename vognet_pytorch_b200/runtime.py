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


class PackedLayout:
    """Byte layout of a batch dict inside ONE flat buffer (every tensor on a 256-byte boundary): lets a whole batch
    travel host -> device as a single copy and be handed to the model as views.  `first` keys come first, in that order
    (the model puts the inputs of its captured forward there, so staging them is one contiguous copy as well)."""
    ALIGN = 256

    def __init__(self, batch, first=()):
        keys = [k for k in first if k in batch] + [k for k in batch if k not in first]
        self.entries, off = [], 0
        for k in keys:
            t = batch[k]
            n = t.numel() * t.element_size()
            self.entries.append((k, off, n, t.dtype, tuple(t.shape)))
            off += -(-n // self.ALIGN) * self.ALIGN
        self.nbytes = off

    def prefix(self, keys):
        """(entries, nbytes) of the leading `keys` - None unless they are exactly the first len(keys) entries"""
        keys = tuple(keys)
        if tuple(e[0] for e in self.entries[:len(keys)]) != keys:
            return None
        ents = tuple(self.entries[:len(keys)])
        end = self.entries[len(keys)][1] if len(self.entries) > len(keys) else self.nbytes
        return ents, end

    def views(self, flat):
        """dict of tensors aliasing the flat uint8 buffer"""
        return {k: flat[o:o + n].view(dt).view(shape) for k, o, n, dt, shape in self.entries}

    def pack_into(self, flat, batch):
        for k, o, n, dt, shape in self.entries:
            t = batch[k]
            if t.dtype != dt or tuple(t.shape) != shape:
                raise ValueError(f'PackedLayout: {k} is {t.dtype} {tuple(t.shape)}, layout has {dt} {shape}')
            flat[o:o + n].view(dt).view(shape).copy_(t)
        return flat


class PackedBatch(dict):
    """A batch dict whose tensors are views of one flat buffer (`.flat`, uint8) described by `.layout`."""

    def __init__(self, flat, layout, slot=None):
        super().__init__(layout.views(flat))
        self.flat, self.layout, self.slot = flat, layout, slot


def pack_host_batch(batch, first=(), layout=None, pin=True):
    """Collate step for the packed path: copy a (CPU) batch dict into one flat, optionally pinned, buffer."""
    layout = layout or PackedLayout(batch, first)
    flat = torch.empty(layout.nbytes, dtype=torch.uint8)
    if pin and torch.cuda.is_available():
        flat = flat.pin_memory()
    layout.pack_into(flat, batch)
    return PackedBatch(flat, layout)


class BatchPrefetcher:
    """Double-buffered host -> device staging: while the model works on batch i (compute stream),
    batch i+1 is copied from pinned host memory on a dedicated copy stream.  ``next()`` returns a
    device batch dict whose copies are ordered before the caller's current stream.

    Host batches may be plain dicts (one copy per tensor, buffers from the caching allocator) or ``PackedBatch``es
    (``pack_host_batch``): then every batch is ONE host->device copy into a ring of ``depth + 1`` preallocated flat
    device buffers and ``next()`` returns a ``PackedBatch`` of views.  A ring slot is overwritten only after the
    event passed to ``release(batch, event)`` for its previous occupant - call it once the step that consumed the
    batch has been enqueued."""

    def __init__(self, host_batches, device, depth=2):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.queue = []
        self.ring, self.consumed, self.n_enq = None, None, 0
        for _ in range(depth):
            self._enqueue()

    def _enqueue(self):
        try:
            hb = next(self.it)
        except StopIteration:
            return
        with torch.cuda.stream(self.copy_stream):
            if isinstance(hb, PackedBatch):
                if self.ring is None or self.ring[0].numel() != hb.layout.nbytes:
                    self.ring = [torch.empty(hb.layout.nbytes, dtype=torch.uint8, device=self.device)
                                 for _ in range(self.depth + 1)]
                    self.consumed = [None] * (self.depth + 1)
                slot = self.n_enq % (self.depth + 1)
                self.n_enq += 1
                if self.consumed[slot] is not None:             # the step that read this slot last has finished
                    self.copy_stream.wait_event(self.consumed[slot])
                self.ring[slot].copy_(hb.flat, non_blocking=True)
                db = PackedBatch(self.ring[slot], hb.layout, slot)
            else:
                db = {k: v.to(self.device, non_blocking=True) for k, v in hb.items()}
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((db, ev))

    def release(self, batch, event):
        """the work recorded by `event` (on the compute stream) is the last reader of a packed `batch`"""
        if isinstance(batch, PackedBatch) and batch.slot is not None and self.consumed is not None:
            self.consumed[batch.slot] = event

    def next(self):
        if not self.queue:
            return None
        db, ev = self.queue.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        if not isinstance(db, PackedBatch):
            for v in db.values():                       # the compute stream now owns these buffers
                v.record_stream(cur)
        self._enqueue()
        return db


class PredictionFetcher:
    """Device -> host read of a step's predictions without stalling the compute stream: the copies run on a dedicated
    stream (ordered after the producing work through an event), into a ring of pinned host buffers.  Tensors that are
    views of one allocation (``ops.select_fwd`` returns boxes / scores / indexs that way) travel as ONE copy.

        f = PredictionFetcher(device)
        f.fetch(i, (boxes, scores, indexs))      # enqueue; returns immediately
        ... enqueue step i + 1 ...
        boxes_h, scores_h, indexs_h = f.get(i)   # blocks until step i's copy has landed
    """

    def __init__(self, device, depth=2):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.slots = [None] * depth          # (signature, pinned flat / list, host views)
        self.done = [None] * depth
        self.nbytes = 0

    @staticmethod
    def _shared(tensors):
        st = tensors[0].untyped_storage()
        if not all(t.untyped_storage().data_ptr() == st.data_ptr() and t.is_contiguous() for t in tensors):
            return False
        return sum(t.numel() * t.element_size() for t in tensors) == st.nbytes()     # the allocation holds nothing else

    def fetch(self, i, tensors):
        tensors = tuple(tensors)
        k = i % self.depth
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        sig = tuple((t.dtype, tuple(t.shape), t.storage_offset()) for t in tensors)
        if self._shared(tensors):
            st = tensors[0].untyped_storage()
            dflat = torch.empty(0, dtype=torch.uint8, device=self.device).set_(st)
            if self.slots[k] is None or self.slots[k][0] != sig:
                hflat = torch.empty(dflat.numel(), dtype=torch.uint8).pin_memory()
                views = []
                for t in tensors:
                    o, n = t.storage_offset() * t.element_size(), t.numel() * t.element_size()
                    views.append(hflat[o:o + n].view(t.dtype).view(t.shape))
                self.slots[k] = (sig, hflat, views)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                self.slots[k][1].copy_(dflat, non_blocking=True)
            dflat.record_stream(self.stream)
            self.nbytes = dflat.numel()
        else:
            if self.slots[k] is None or self.slots[k][0] != sig:
                views = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in tensors]
                self.slots[k] = (sig, None, views)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                for h, t in zip(self.slots[k][2], tensors):
                    h.copy_(t, non_blocking=True)
            for t in tensors:
                t.record_stream(self.stream)
            self.nbytes = sum(t.numel() * t.element_size() for t in tensors)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        self.done[k] = ev
        return ready

    def get(self, i):
        k = i % self.depth
        self.done[k].synchronize()
        return self.slots[k][2]
