"""The data-parallel training step as ONE call (BASELINE.json configs[4]; utils/trn_utils.py:497-505,
code/main_dist.py:55,75-80):

    out = mdl(batch); loss = loss_fn(out, batch)['loss']; optimizer.zero_grad(); loss.backward(); optimizer.step()

``FusedTrainStep(mdl, loss_fn, FlatAdam)`` runs exactly that arithmetic, but
  * the backward kernels accumulate straight into the optimizer's flat gradient buffer (no per-parameter autograd
    accumulation pass over 45 M parameters),
  * the gradient mean over the ranks is TWO NCCL all-reduces over contiguous slices of that buffer - the language side
    (LSTM + embeddings = 82 % of the bytes) as soon as its gradients exist, which is before the object transformer's
    backward runs, and the rest at the end - instead of DistributedDataParallel's ~25 MB buckets,
  * Adam is one kernel over the flat buffers (``vog_adam_step``) with the 1/world of the mean folded in.
The plain ``mdl(batch)`` / ``loss.backward()`` protocol of the reference trainer keeps working (``training._TrainFn``);
this class is the fast path the benchmark times."""
import torch
import torch.distributed as dist

from . import training
from .optim import FlatAdam
from .training import GradSink


class FusedTrainStep:
    def __init__(self, mdl, loss_fn, optimizer, group=None, overlap=True):
        if not isinstance(optimizer, FlatAdam):
            raise TypeError('FusedTrainStep needs a vognet_pytorch_b200.optim.FlatAdam optimizer (flat gradient buffer)')
        self.mdl, self.loss_fn, self.opt, self.group, self.overlap = mdl, loss_fn, optimizer, group, overlap
        name_of = {id(p): n for n, p in mdl.named_parameters()}
        self.views, lang_end = {}, 0
        for p, o in zip(optimizer.params, optimizer.offsets):
            n = name_of[id(p)]
            self.views[n] = optimizer.flat_grad[o:o + p.numel()].view_as(p)
            if n.startswith(('lstm_encoder.', 'lstm_out_feat_proj.', 'srl_arg_words_out_enc.')):
                lang_end = max(lang_end, o + (p.numel() + 3) // 4 * 4)
        # the language-side parameters are registered first (code/mdl_vog.py:160-193 builds the language model
        # first), so their gradients are one contiguous prefix of the flat buffer
        first_other = min((o for p, o in zip(optimizer.params, optimizer.offsets)
                           if not name_of[id(p)].startswith(('lstm_encoder.', 'lstm_out_feat_proj.',
                                                             'srl_arg_words_out_enc.'))), default=lang_end)
        self.lang_end = lang_end if lang_end <= first_other else 0
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.comm = self.world > 1
        self.allreduce_bytes = optimizer.numel * 4

    def _reduce(self, lo, hi):
        if hi > lo:
            return dist.all_reduce(self.opt.flat_grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return None

    def __call__(self, batch):
        mdl, opt = self.mdl, self.opt
        if not mdl.training:
            raise RuntimeError('FusedTrainStep: call model.train() first')
        training._check_supported(mdl)
        tc = mdl.compute != 'fp32x'
        if tc:
            from . import training_tc
            logits, tape = training_tc.forward_train_tc(mdl, batch)
        else:
            logits, tape = training.forward_train_f32(mdl, batch)
        lg = logits.detach().requires_grad_(True)
        loss = self.loss_fn({'mdl_outs': lg}, batch)['loss']
        loss.backward()                                   # the loss kernels only: d loss / d logits
        opt.flat_grad.zero_()
        sink = GradSink(mdl.named_parameters(), views=self.views)
        works = []

        def lang_done():
            if self.comm and self.overlap and self.lang_end:
                works.append(self._reduce(0, self.lang_end))
        if tc:
            training_tc.backward_train_tc(mdl, tape, lg.grad, sink=sink, on_lang_done=lang_done)
        else:
            training.backward_train_f32(mdl, tape, lg.grad, sink=sink, on_lang_done=lang_done)
        if self.comm:
            works.append(self._reduce(self.lang_end if works else 0, opt.numel))
            for w in works:
                if w is not None:
                    w.wait()                              # orders the current stream after the collective
            opt._grad_scale = 1.0 / self.world
        opt.step()
        return loss.detach()


def bench_train_step(workload, compute, dev, world, rank, flush, barrier, max_ranks, steps, warmup=3, overlap=True):
    """Timed training steps on one synthetic batch per rank (weak scaling: per-GPU batch fixed).  -> dict for bench.py."""
    import vognet_pytorch_b200 as vb
    from . import _lib, synth
    L = _lib.lib()
    w, batch = synth.workload(workload, seed=1 + rank)
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    cfg, comm = synth.default_cfg(w['conc_type']), synth.default_comm(w['nppf'])
    sel = vb.get_mdl_loss_eval(cfg)
    mdl = sel['mdl'](cfg, comm)
    mdl.load_state_dict(synth.make_state_dict(), strict=True)
    mdl = mdl.to(dev).set_compute(compute).train()
    loss_fn = sel['loss'](cfg, comm)
    opt = FlatAdam(mdl.parameters(), lr=1e-4, betas=(0.9, 0.99))
    step = FusedTrainStep(mdl, loss_fn, opt, overlap=overlap)
    dinp = {k: v.to(dev) for k, v in inp.items()}
    B = w['B']
    torch.manual_seed(1234 + rank)
    losses = []
    for _ in range(max(warmup, 1)):
        losses.append(step(dinp))
    barrier()
    torch.cuda.reset_peak_memory_stats(dev)
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    n0 = L.vog_launch_count()
    barrier()
    for i in range(steps):
        flush.zero_()
        e0[i].record()
        losses.append(step(dinp))
        e1[i].record()
    barrier()
    launches = (L.vog_launch_count() - n0) // steps
    per = [a.elapsed_time(b) for a, b in zip(e0, e1)]
    t_ms = max_ranks(sum(per))
    out = {'metric': f'VOGNet training-step queries/sec ({workload})', 'value': world * B * steps / (t_ms / 1e3),
           'unit': 'queries/s', 'steps_per_s': steps / (t_ms / 1e3), 'ms_per_step': t_ms / steps, 'steps': steps,
           'warmup': max(warmup, 1), 'n_gpus': world, 'per_gpu_batch': B, 'global_batch': world * B, 'compute': compute,
           'higher_is_better': True, 'scaling': 'weak', 'data': 'synthetic',
           'dropout': bool(mdl.train_dropout), 'optimizer': 'FlatAdam(0.9, 0.99), one kernel over 45.25 M parameters',
           'gpu_launches': int(launches), 'allreduce_bytes': step.allreduce_bytes if world > 1 else 0,
           'allreduce': ('2 NCCL all-reduces over contiguous slices of the flat fp32 gradient (language side first, '
                         'overlapped with the object transformer backward)') if world > 1 else 'none (1 GPU)',
           'peak_mem_gib': torch.cuda.max_memory_allocated(dev) / 2 ** 30,
           'loss_first': float(losses[0]), 'loss_last': float(losses[-1]),
           'ms_per_step_median': sorted(per)[len(per) // 2], 'ms_per_step_min': min(per),
           'config': {'workload': workload, 'conc_type': w['conc_type'], 'per_gpu_batch': B, 'global_batch': world * B,
                      'ncmp': w['ncmp'], 'nfrm': 10, 'nppf': w['nppf'], 'compute': compute,
                      'l2': 'flushed (256 MB memset) between timed iterations',
                      'parallelism': f'dp{world} (queries sharded, flat gradient all-reduce)'},
           'note': 'forward (dropout on) + LossB + backward + gradient all-reduce + Adam on the same resident batch'}
    if world > 1:
        # the same steps without the collective: exposed communication = difference of the two step times
        step.comm = False
        for _ in range(2):
            step(dinp)
        barrier()
        ts = []
        for i in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(dinp); b.record()
            ts.append((a, b))
        barrier()
        t_nc = max_ranks(sum(a.elapsed_time(b) for a, b in ts))
        out['ms_per_step_no_allreduce'] = t_nc / steps
        out['exposed_comm_us'] = max(0.0, (t_ms - t_nc) / steps) * 1e3
        # the collective alone (both slices back to back, nothing else on the GPU)
        g = opt.flat_grad
        for _ in range(2):
            dist.all_reduce(g)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dist.all_reduce(g[:step.lang_end]) if step.lang_end else None
            dist.all_reduce(g[step.lang_end:])
        b.record()
        barrier()
        ar_ms = max_ranks(a.elapsed_time(b)) / 5
        out['allreduce_alone_us'] = ar_ms * 1e3
        out['allreduce_busbw_gbs'] = step.allreduce_bytes * 2 * (world - 1) / world / (ar_ms * 1e-3) / 1e9
        step.comm = True
    return out
