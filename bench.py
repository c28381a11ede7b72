#!/usr/bin/env python
"""bench.py - VOGNet forward queries/sec on B200 (driver contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload spat_gt5|spat_p100|temp_gt5|temp_p100]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU oracle port on the host cores, same workload

One "step" = one forward of the fusion hot path (language LSTM, encoders, obj_tx, mul_tx, lin2,
masks) over one synthetic batch of B queries per GPU + the box selection.  `value` = queries/s
with inputs resident in HBM; `e2e` = the same through the public nn.Module / evaluator API with the
batch starting in pinned HOST memory and predictions read back to the host every step.

The default workload is BASELINE.json configs[1] (spat/gt5, bs=4 per GPU, fp32 configuration ->
compute 'tf32': tf32 tcgen05 GEMMs, bf16 attention operands, fp32 everything else), the one the
metric is quoted on.  The roofline object is reported for the dominant kernel of the workload;
`roofline_seq4000` repeats it for the fused attention at the north-star shape (spat/p100, N=4000 /
2000) so the tensor-pipe target can be read from the same line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'VOGNet fwd queries/sec'
COMPUTE = {'spat_gt5': 'tf32', 'temp_gt5': 'tf32', 'spat_p100': 'bf16', 'temp_p100': 'bf16', 'cpu_ref': 'tf32'}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


_SAMPLER_SRC = r"""
import sys, time
idx = int(sys.argv[1])
try:
    import pynvml as nv
    nv.nvmlInit()
    h = nv.nvmlDeviceGetHandleByIndex(idx)
    mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
    get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
    print('max', mx, flush=True)
    while True:
        print(repr(time.time()), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(get(h)), flush=True)
        time.sleep(0.002)
except Exception as e:
    print('err', repr(e), flush=True)
"""


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region by a separate PROCESS polling NVML every
    ~2 ms (a thread would fight the launching thread for the GIL, and `nvidia-smi -lms` is too coarse for a
    timed region of a few tens of milliseconds); samples are time-stamped and filtered to the window
    [mark_begin(), mark_end()]."""
    BITS = {'sw_power_cap': 0x4, 'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40}

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def start(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = self.index
        if vis:
            try:
                phys = int(vis.split(',')[self.index])
            except Exception:
                phys = self.index
        try:
            self.proc = subprocess.Popen([sys.executable, '-c', _SAMPLER_SRC, str(phys)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['sampler unavailable'], 'samples': 0}
        time.sleep(0.01)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=1)
        mx, sm, mask, err = None, [], 0, None
        for ln in self.lines:
            f = ln.split()
            if not f:
                continue
            if f[0] == 'max':
                mx = float(f[1])
            elif f[0] == 'err':
                err = ln.strip()
            else:
                try:
                    t, c, m = float(f[0]), float(f[1]), int(f[2])
                except Exception:
                    continue
                if (self.t0 is None or t >= self.t0) and (self.t1 is None or t <= self.t1):
                    sm.append(c)
                    mask |= m
        reasons = sorted(n for n, b in self.BITS.items() if mask & b)
        out = {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx, 'reasons': reasons,
               'samples': len(sm), 'source': 'nvml polled every 2 ms by a side process, samples inside the timed window'}
        if err:
            out['error'] = err
        return out


# ---------------------------------------------------------------------------------------------
def flops_attn(Bt, N, d):
    """QK^T + PV of the fused attention kernel (multiply-add = 2): SURVEY.md section 8d."""
    return Bt * 4.0 * N * N * d


def workload_shapes(w):
    B, ncmp, nppf = w['B'], w['ncmp'], w['nppf']
    P = ncmp * 10 * nppf
    if w['conc_type'] == 'spat':
        nfrm, nppf2 = 10, ncmp * nppf
    else:
        nfrm, nppf2 = ncmp * 10, nppf
    return dict(P=P, obj=(B, P, 512), mul=(B * nfrm, 5 * nppf2, 768))


def flops_query(w):
    s = workload_shapes(w)
    P = s['P']

    def layer(Bt, N, d):
        return Bt * (6.0 * N * d * d + 4.0 * N * N * d) + Bt * (2.0 * N * d * d + 4.0 * N * d * (d / 2))
    B = w['B']
    tot = layer(*s['obj']) + layer(*s['mul'])
    tot += B * (2.0 * P * 2048 * 256 + 2.0 * 40 * 3072 * 256 + 2.0 * 5 * P * (768 * 256 + 256))
    return tot / B


def _reference_step_fn(w, sd, device):
    """One forward + box selection of the CPU/GPU baseline: the UNMODIFIED reference (from /root/reference or the
    baseline/_ref copy build() installs) when present, else the oracle port.  -> (step(batch), kind)."""
    import torch
    from oracle import ref_harness as rh
    if rh.reference_available():
        mdl = rh.build_reference_model(w['conc_type'], w['nppf'], sd).to(device)
        ev = rh.build_reference_evaluator(w['conc_type'], w['nppf'], w['ncmp'])

        def step(batch):
            b = dict(batch)                      # the reference edits this key in place (code/mdl_vog.py:80-82)
            b['srl_arg_word_mask'] = batch['srl_arg_word_mask'].clone()
            with torch.no_grad():
                out = mdl(b)
                return ev.get_out_results_boxes(out, b)
        return step, 'reference'
    from oracle import vog_oracle as vo
    sdd = {k: v.to(device) for k, v in sd.items()}

    def step(batch):
        with torch.no_grad():
            out = vo.vog_forward(sdd, batch, w['conc_type'], w['nppf'])
            return vo.select_boxes(out['mdl_outs_eval'], batch['pad_proposals'], w['conc_type'], w['ncmp'], w['nppf'])
    return step, 'port'


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (the unmodified
    reference from baseline/_ref, installed by build(); the oracle port only if that copy is missing)."""
    import torch
    from vognet_pytorch_b200 import synth
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    w, batch = synth.workload(args.workload)
    sd = synth.make_state_dict()
    B = w['B']
    sample = f'whole per-GPU batch B={B} on rank 0 only'
    if w['nppf'] >= 100:           # ~2 s per query on 8 cores: bound the sample to one query
        batch = {k: v[:1].clone() for k, v in batch.items()}
        B, sample = 1, 'first query of the batch (B=1) on rank 0 only'
    step, kind = _reference_step_fn(w, sd, 'cpu')
    for _ in range(args.warmup):
        step(batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(batch)
    dt = (time.perf_counter() - t0) / args.steps
    v = B / dt
    line = {'impl': 'reference', 'metric': f'{METRIC} ({args.workload})', 'value': v, 'unit': 'queries/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            # same workload description as the b200 arm's line; the CPU arm always runs ONE per-GPU batch on ONE
            # host (rank 0), whatever --gpus says: at N > 1 the b200 arm processes N x as many queries per step
            'config': {'workload': args.workload, 'conc_type': w['conc_type'], 'per_gpu_batch': w['B'],
                       'global_batch': w['B'], 'ncmp': w['ncmp'], 'nfrm': 10, 'nppf': w['nppf'],
                       'obj_attn': list(workload_shapes(w)['obj']), 'mul_attn': list(workload_shapes(w)['mul']),
                       'compute': 'fp32 (torch CPU ops, eval mode, no_grad)',
                       'parallelism': 'host cores of rank 0', 'sample_queries_per_step': B},
            'reference_n_queries': B * args.steps, 'reference_ranks': 1,
            'cpu_baseline': {'value': v, 'unit': 'queries/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': v, 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def train_step_object(dev, world, rank, flush, barrier, max_ranks, steps):
    """BASELINE.json configs[4]: spat/p100 TRAINING step, bs=4 per GPU (bs=32 at 8 GPUs): forward + LossB_SPAT +
    backward + ONE flat NCCL gradient all-reduce + fused Adam (utils/trn_utils.py:497-505, code/main_dist.py:55,75-80)."""
    from vognet_pytorch_b200 import train_step
    return train_step.bench_train_step('spat_p100', 'bf16', dev, world, rank, flush, barrier, max_ranks, steps)


def gpu_eager_train_object(workload, dev, flush, steps):
    """The in-box GPU comparator of the TRAINING step: the unmodified reference (train mode, dropout on) + its own
    LossB + torch.optim.Adam(0.9, 0.99) under torch eager / autograd on the same B200 - utils/trn_utils.py:497-505,
    code/main_dist.py:55 - same synthetic batch, inputs resident, L2 flushed between steps, CUDA events."""
    import torch
    from oracle import ref_harness as rh
    from vognet_pytorch_b200 import synth
    if not rh.reference_available():
        return {'unavailable': 'baseline/_ref is not installed'}
    w, batch = synth.workload(workload)
    inp = dict(batch)
    inp.update(synth.make_loss_inputs(batch, **w))
    mdl = rh.build_reference_model(w['conc_type'], w['nppf'], synth.make_state_dict()).to(dev).train()
    loss_fn = rh.build_reference_loss(w['conc_type'], w['nppf']).to(dev)
    opt = torch.optim.Adam(mdl.parameters(), lr=1e-4, betas=(0.9, 0.99))
    b = {k: v.to(dev) for k, v in inp.items()}

    def step():
        bb = dict(b)
        bb['srl_arg_word_mask'] = b['srl_arg_word_mask'].clone()     # edited in place by the reference
        opt.zero_grad()
        loss = loss_fn(mdl(bb), bb)['loss']
        loss.backward()
        opt.step()
        return loss
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(dev)
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del mdl, opt
    torch.cuda.empty_cache()
    return {'value': w['B'] / (ms * 1e-3), 'unit': 'queries/s', 'ms_per_step': ms, 'steps': steps, 'kind': 'reference',
            'dtype': 'f32', 'torch': torch.__version__, 'peak_mem_gib': peak,
            'note': 'unmodified reference forward (train mode) + LossB + autograd backward + torch.optim.Adam under '
                    'torch eager on the same B200 (library kernels only)'}


def run_train(args, dev, world, rank, sampler, flush, barrier, max_ranks):
    from vognet_pytorch_b200 import train_step
    compute = args.compute or ('bf16' if COMPUTE[args.workload] != 'fp32x' else 'fp32x')
    sampler.mark_begin()
    obj = train_step.bench_train_step(args.workload, compute, dev, world, rank, flush, barrier, max_ranks, args.steps,
                                    warmup=args.warmup)
    sampler.mark_end()
    clocks = sampler.stop()
    if rank == 0:
        obj['clocks'] = clocks
        if world == 1 and not args.no_extras:
            try:
                obj['gpu_eager_baseline'] = gpu_eager_train_object(args.workload, dev, flush, 5)
            except Exception as e:                             # never let a baseline leg take the line down
                obj['gpu_eager_baseline'] = {'unavailable': repr(e)[:300]}
        print(json.dumps(obj), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def gpu_eager_object(w, batch, sd, dev, flush, steps):
    """The in-box GPU comparator: the unmodified reference (or the oracle port when baseline/_ref is missing) on
    cuda under torch eager - ATen / cuBLAS / cuDNN kernels, fp32 - same batch, inputs resident, L2 flushed
    between steps, CUDA events."""
    import torch
    step, kind = _reference_step_fn(w, sd, dev)
    b = {k: v.to(dev) for k, v in batch.items()}
    for _ in range(3):
        step(b)
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(b); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    peak = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    del step
    torch.cuda.empty_cache()
    return {'value': w['B'] / (ms * 1e-3), 'unit': 'queries/s', 'ms_per_step': ms, 'steps': steps, 'kind': kind,
            'dtype': 'f32', 'torch': torch.__version__, 'tf32_matmul': bool(torch.backends.cuda.matmul.allow_tf32),
            'peak_mem_gib': peak,
            'note': 'reference forward + evaluator selection under torch eager on the same B200 (library kernels only)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--workload', default='spat_gt5')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--compute', default=None, help="override: fp32x | tf32 | bf16")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-seq4000', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the sub-objects (north_star_forward, '
                    'gpu_eager_baseline, value_fp32x, temp_gt5, train_step)')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of CUDA graphs')
    ap.add_argument('--train', action='store_true', help='time the TRAINING step (forward + loss + backward + '
                    'flat gradient all-reduce + fused Adam) instead of the forward')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        if args.steps == 100:
            args.steps = 5
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import vognet_pytorch_b200 as vb
    from vognet_pytorch_b200 import _lib, ops, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        # a rank that dies must not leave the others (and the box) waiting for the default 10-minute watchdog
        dist.init_process_group('nccl', device_id=torch.device('cuda', local), timeout=datetime.timedelta(seconds=180))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    L = _lib.lib()
    sampler = ClockSampler(local)
    sampler.start()              # side process: up and polling NVML long before the timed region starts
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2
    sd_cpu = synth.make_state_dict()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def build(workload, compute, graph=True):
        """model + evaluator + (host, resident) batches of one workload; every rank its own shard of queries"""
        w_, batch_ = synth.workload(workload, seed=1 + rank)
        cfg, comm = synth.default_cfg(w_['conc_type']), synth.default_comm(w_['nppf'])
        sel = vb.get_mdl_loss_eval(cfg)
        m = sel['mdl'](cfg, comm)
        m.load_state_dict(sd_cpu, strict=True)
        m = m.to(dev).eval().set_compute(compute)
        m.use_cuda_graph = graph and compute != 'fp32x'
        e = sel['eval'](cfg, comm, dev)
        res = {k: v.to(dev) for k, v in batch_.items()}
        if m.use_cuda_graph:
            # "inputs already resident in HBM": the resident batch lives in the captured forward's own input
            # tensors, so the timed step is the graph replay alone (no staging copies)
            with torch.no_grad():
                res = m.graph_input_buffers(res)
        return dict(w=w_, batch=batch_, mdl=m, ev=e, resident=res, sel=sel, cfg=cfg, comm=comm)

    def fwd_step(ctx, b):
        out = ctx['mdl'](b)
        return out, ctx['ev'].get_out_results_boxes(out, b)

    def time_resident(ctx, steps, warmup, local=False):
        """K steps on the resident batch, L2 flushed between steps, one CUDA-event pair per step; max over ranks.
        local=True: a rank-0-only sub-object - NO collective (the other ranks have left by then)."""
        sync = torch.cuda.synchronize if local else barrier
        for _ in range(warmup):
            fwd_step(ctx, ctx['resident'])
        sync()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        n0 = L.vog_launch_count()
        sync()
        for i in range(steps):
            flush.zero_()
            e0[i].record()
            fwd_step(ctx, ctx['resident'])
            e1[i].record()
        sync()
        launches = (L.vog_launch_count() - n0) // steps          # eager launches of libvog_b200 per step
        if ctx['mdl'].use_cuda_graph:                             # + the kernels captured in the replayed graph
            launches += int(getattr(ctx['mdl'], 'graph_launches', 0))
        per = [a.elapsed_time(b) for a, b in zip(e0, e1)]
        return (sum(per) if local else max_ranks(sum(per))), per, launches

    if args.train:
        return run_train(args, dev, world, rank, sampler, flush, barrier, max_ranks)

    compute = args.compute or COMPUTE[args.workload]
    ctx = build(args.workload, compute, not args.no_graph)
    w, batch, mdl, ev, resident = ctx['w'], ctx['batch'], ctx['mdl'], ctx['ev'], ctx['resident']
    B = w['B']
    host = {k: v.pin_memory() for k, v in batch.items()}

    def step(b):
        return fwd_step(ctx, b)

    # ---- kernel-resident timing: K steps, L2 flushed between steps, CUDA events per step ---------
    for _ in range(args.warmup):
        step(resident)
    barrier()
    sampler.mark_begin()
    t_ms, per_step, launches = time_resident(ctx, args.steps, 0)
    sampler.mark_end()
    clocks = sampler.stop()
    value = world * B * args.steps / (t_ms / 1e3)

    # ---- end to end: pinned host batch -> H2D -> forward + selection -> D2H of the predictions ----
    h2d = None          # set below: bytes of the packed batch that actually cross PCIe every step
    d2h = 0

    from vognet_pytorch_b200.runtime import BatchPrefetcher, pack_host_batch
    # every step's batch starts in pinned host memory; the copy of step i+1 overlaps the compute of
    # step i on a copy stream (what a pinned-memory DataLoader feeding the reference does, too)
    # the collate step of the packed path: the whole batch in ONE pinned buffer (the model's graph inputs first), so
    # every step's host->device transfer is a single copy and its staging into the captured forward another one
    graph_keys = tuple(k for k in getattr(mdl, '_GRAPH_KEYS', ()) if k in host)
    host_packed = pack_host_batch(host, first=graph_keys)
    h2d = int(host_packed.layout.nbytes)
    pre = BatchPrefetcher((host_packed for _ in range(args.warmup + 6 * args.steps)), dev)
    # Two steps are in flight: while step i runs on the GPU the host stages step i+1 and then collects the
    # predictions of step i from pinned memory (event wait).  Every step still pays its own H2D copy (from
    # pinned host memory, on the copy stream) and its own D2H read of boxes / scores / indexs.
    from vognet_pytorch_b200.runtime import PredictionFetcher
    fetcher = PredictionFetcher(dev)        # D2H on its own stream: the next step's forward does not queue behind it

    def launch(i):
        nonlocal d2h
        b = pre.next()
        out, s = step(b)
        ev_ = fetcher.fetch(i, (s['boxes'], s['scores'], s['indexs']))     # event: the step's compute is enqueued
        pre.release(b, ev_)                                # the batch's ring slot may be refilled after this step
        d2h = fetcher.nbytes

    def collect(i):
        return fetcher.get(i)                              # the step's predictions are on the host

    def e2e_run(n):
        for i in range(n):
            launch(i)
            if i:
                collect(i - 1)
        collect(n - 1)
    e2e_run(args.warmup)
    e2e_run(args.steps)          # one untimed block: pinned buffers, allocator pools and clocks settle
    # K steps per block, every block bracketed by a barrier + synchronize; a 20-step block lasts ~15 ms, so one host
    # hiccup (page fault, scheduler) moves it by 10-20 %: the block is repeated and the MEDIAN block is reported
    e2e_blocks = []
    for _ in range(5):
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_blocks.append(world * B * args.steps / te.item())
    e2e = statistics.median(e2e_blocks)

    # ---- sub-objects that every rank takes part in (weak scaling, same rules as the headline) --------------
    extras = {}
    if not args.no_extras:
        sub_steps = max(5, min(args.steps, 20))
        if args.workload != 'temp_gt5':
            # BASELINE.json configs[3]: temp/gt5, bs=4 per GPU (bs=8 global on 2 GPUs)
            c2 = build('temp_gt5', COMPUTE['temp_gt5'])
            t2, _, l2 = time_resident(c2, sub_steps, 3)
            extras['temp_gt5'] = {'value': world * c2['w']['B'] * sub_steps / (t2 / 1e3), 'unit': 'queries/s',
                                  'ms_per_step': t2 / sub_steps, 'steps': sub_steps, 'n_gpus': world,
                                  'global_batch': world * c2['w']['B'], 'compute': COMPUTE['temp_gt5'],
                                  'gpu_launches': int(l2), 'note': 'BASELINE.json configs[3] (bs=8 global at 2 GPUs); '
                                  'resident inputs, CUDA graph, L2 flushed between steps'}
            del c2
        try:
            extras['train_step'] = train_step_object(dev, world, rank, flush, barrier, max_ranks, sub_steps)
        except (NotImplementedError, ImportError, AttributeError) as e:   # backward not built for this configuration
            extras['train_step'] = {'unavailable': str(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank-0 sub-objects ---------------------------------------------------------------------------------
    pk, pk_kind = peaks()
    if not args.no_extras:
        sub_steps = max(5, min(args.steps, 20))
        # (a) the exact-fp32 CUDA-core path on the headline workload (BASELINE config 2 says fp32: this is the
        #     number with fp32 arithmetic end to end; the headline runs tf32 GEMMs + bf16 attention operands)
        if compute != 'fp32x':
            cx = build(args.workload, 'fp32x', False)
            tx, _, lx = time_resident(cx, sub_steps, 3, local=True)
            extras['value_fp32x'] = {'value': B * sub_steps / (tx / 1e3), 'unit': 'queries/s', 'ms_per_step': tx / sub_steps,
                                     'steps': sub_steps, 'gpu_launches': int(lx),
                                     'note': "compute='fp32x': every product and sum in IEEE fp32 on CUDA cores, eager launches"}
            del cx
        # (b) the in-box GPU comparator: the UNMODIFIED reference under torch eager on this B200,
        #     same batch, resident inputs, CUDA events
        try:
            extras['gpu_eager_baseline'] = gpu_eager_object(w, batch, sd_cpu, dev, flush, sub_steps)
        except Exception as e:                                 # never let a baseline leg take the line down
            extras['gpu_eager_baseline'] = {'unavailable': repr(e)[:300]}
        # (c) the north-star shape: full spat/p100 bs=4 forward (obj N=4000, mul N=2000), resident, CUDA graph
        if args.workload != 'spat_p100' and world == 1:
            cn = build('spat_p100', 'bf16')
            tn, _, ln = time_resident(cn, sub_steps, 3, local=True)
            fq = flops_query(cn['w']) * cn['w']['B']
            tf = fq / (tn / sub_steps * 1e-3) / 1e12
            ns = {'workload': 'spat_p100', 'per_gpu_batch': cn['w']['B'], 'value': cn['w']['B'] * sub_steps / (tn / 1e3),
                  'unit': 'queries/s', 'ms_per_step': tn / sub_steps, 'steps': sub_steps, 'compute': 'bf16',
                  'gpu_launches': int(ln), 'tflops_algorithmic': tf,
                  'frac_of_sustained_bf16': tf / pk['bf16_tflops_sustained'],
                  'note': 'BASELINE.json configs[2]; algorithmic FLOPs of SURVEY 8d (dense QKV formula) / measured time'}
            try:
                ge = gpu_eager_object(cn['w'], cn['batch'], sd_cpu, dev, flush, 5)
                ns['gpu_eager_baseline'] = ge
            except Exception as e:
                ns['gpu_eager_baseline'] = {'unavailable': repr(e)[:300]}
            extras['north_star_forward'] = ns
            del cn
        if world == 1 and isinstance(extras.get('train_step'), dict) and 'value' in extras['train_step']:
            try:
                extras['train_step']['gpu_eager_baseline'] = gpu_eager_train_object('spat_p100', dev, flush, 3)
            except Exception as e:
                extras['train_step']['gpu_eager_baseline'] = {'unavailable': repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (fused attention), timed alone with CUDA events ---------

    def traffic_of(kernel, shape_key):
        """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
        (profiles/r2/traffic.json, else profiles/r1/traffic.json); None when this shape was not captured."""
        for rnd in ('r2', 'r1'):                       # the newest capture of this kernel / shape wins
            try:
                t = json.load(open(os.path.join(ROOT, 'profiles', rnd, 'traffic.json')))
            except Exception:
                continue
            v = t.get(kernel, {}).get(shape_key)
            if v is not None:
                return v
        return None

    def attn_roofline(shape, n_iter=10):
        Bt, N, d = shape
        hd = ops.chunk_sizes(d, 3)
        dhp = ops.round_up(max(hd), 64)
        g = torch.Generator(device='cpu').manual_seed(0)
        q = (torch.rand(Bt, 3, N, dhp, generator=g) - 0.5).bfloat16().to(dev)
        k = (torch.rand(Bt, 3, N, dhp, generator=g) - 0.5).bfloat16().to(dev)
        vt = (torch.rand(Bt, 3, N, dhp, generator=g) - 0.5).bfloat16().to(dev)
        nbox = N // 5 if N % 5 == 0 else N
        a = torch.rand(Bt * nbox, 3, generator=g).to(dev)
        bpe = torch.zeros(3, device=dev)
        out = torch.empty(Bt * N, 3 * dhp, device=dev, dtype=torch.bfloat16)

        def run():
            ops.tc_attn_fwd(q, k, vt, N, hd, 1.0 / d ** 0.5, out=out, bias_mode=ops.BIAS_RANK1, a=a,
                            nbox=nbox, bpe=bpe)
        for _ in range(3):
            run()
        ts = []
        for _ in range(n_iter):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        fl = flops_attn(Bt, N, d)
        ach = fl / (ms * 1e-3) / 1e12
        return {'bound': 'tensor', 'achieved': ach, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s',
                'frac': ach / pk['bf16_tflops'], 'traffic': traffic_of('tc_attn2_kernel', f'Bt{Bt}_N{N}_d{d}'),
                'kernel': 'tc_attn2_kernel',
                'shape': {'Bt': Bt, 'N': N, 'd_model': d, 'heads': 3}, 'ms_per_launch': ms,
                'flops_per_launch': fl, 'peak_source': f'{pk_kind} bf16 burst (kernel timed alone)'}

    def lstm_roofline(n_iter=10):
        """The recurrence kernel of one LSTM layer (the dominant kernel of the gt5 workloads: 2 launches =
        about half of the step).  HBM-side roofline: its algorithmic traffic is W_hh once (the weight-resident
        kernel keeps it on chip for all timesteps) + the input projections in + the hidden states out."""
        T, Bq, Hh = 20, B, 1024
        lp = ops.LP_BF16 if compute == 'bf16' else ops.LP_TF32
        g = torch.Generator(device='cpu').manual_seed(0)
        gx = (torch.rand(T * Bq, 8 * Hh, generator=g) - 0.5).to(dev)
        whh = ((torch.rand(2, 4 * Hh, Hh, generator=g) - 0.5) / 32).to(dev)
        lens = batch['srl_arg_word_mask_len'].reshape(-1)[:Bq].to(dev)

        def run():
            ops.lstm_layer_fwd(gx, whh, lens, T, Bq, lp)
        for _ in range(3):
            run()
        ts = []
        for _ in range(n_iter):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        steps_t = int(lens.max().item())
        by = whh.numel() * 4 + steps_t * Bq * 8 * Hh * 4 + T * Bq * 2 * Hh * (2 if lp == ops.LP_BF16 else 4)
        ach = by / (ms * 1e-3) / 1e9
        # 3-4 sentences per launch run the two-units-per-warp variant of the weight-resident kernel (csrc/lstm_rec.cu)
        kname = 'lstm_rec_pair_kernel' if 3 <= Bq <= 4 else 'lstm_rec_resident_kernel'
        return {'bound': 'hbm', 'achieved': ach, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / pk['hbm_gbs'],
                'traffic': traffic_of(kname, f'T{T}_B{Bq}_H{Hh}'),
                'kernel': kname, 'shape': {'T': T, 'Bq': Bq, 'H': Hh, 'steps': steps_t},
                'ms_per_launch': ms, 'bytes_per_launch': by, 'launches_per_step': 2,
                'peak_source': f'{pk_kind} HBM copy bandwidth',
                'note': 'latency-bound: one cross-SM h_t exchange per timestep and direction (18 dependent steps); '
                        'the bytes are W_hh read once + gate pre-activations in + hidden states out'}

    shapes = workload_shapes(w)
    roof_attn = attn_roofline(shapes['mul'])
    roof = lstm_roofline() if w['nppf'] < 100 else roof_attn
    line = {
        'metric': f'{METRIC} ({args.workload})', 'value': value, 'unit': 'queries/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'tf32': 'tf32', 'bf16': 'bf16', 'fp32x': 'f32'}[compute],
        'data': 'synthetic',
        'config': {'workload': args.workload, 'conc_type': w['conc_type'], 'per_gpu_batch': B,
                   'global_batch': B * world, 'ncmp': w['ncmp'], 'nfrm': 10, 'nppf': w['nppf'],
                   'obj_attn': list(shapes['obj']), 'mul_attn': list(shapes['mul']), 'compute': compute,
                   'compute_detail': {'tf32': 'tf32 tcgen05 GEMMs, bf16 attention operands, fp32 accumulate / softmax / LayerNorm / residual stream',
                                      'bf16': 'bf16 tcgen05 operands, fp32 accumulate / softmax / LayerNorm / residual stream',
                                      'fp32x': 'exact fp32 CUDA-core path'}[compute],
                   'parallelism': f'dp{world} (queries sharded, no data-path collective)',
                   'l2': 'flushed (256 MB memset) between timed iterations',
                   'gflop_per_query_algorithmic': flops_query(w) / 1e9},
        'clocks': clocks,
        'e2e': {'value': e2e, 'unit': 'queries/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'blocks': [round(x, 1) for x in e2e_blocks],
                'note': 'public nn.Module + evaluator API, pinned host batch in / predictions out every step, '
                        'two steps in flight (H2D of step i+1 and D2H of step i-1 overlap the compute of step i), the batch '
                        'travels as one packed pinned buffer (runtime.pack_host_batch); '
                        'median of 5 blocks of `steps` steps (max over ranks per block)'},
        'gpu_launches': int(launches),
        'roofline': roof,
        'roofline_attention': roof_attn,
    }
    if not args.no_seq4000 and world == 1:
        p100 = dict(synth.WORKLOADS['spat_p100'])
        s4 = workload_shapes(p100)
        r_obj, r_mul = attn_roofline(s4['obj'], 5), attn_roofline(s4['mul'], 5)
        fl = r_obj['flops_per_launch'] + r_mul['flops_per_launch']
        ms = r_obj['ms_per_launch'] + r_mul['ms_per_launch']
        line['roofline_seq4000'] = {'bound': 'tensor', 'achieved': fl / ms / 1e9, 'peak': pk['bf16_tflops'],
                                    'unit': 'TFLOP/s', 'frac': fl / ms / 1e9 / pk['bf16_tflops'],
                                    'obj_tx': r_obj, 'mul_tx': r_mul,
                                    'note': 'fused obj_tx+mul_tx attention at spat/p100 bs=4 (N=4000 / 2000)'}
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        torch.set_num_threads(cores)
        cb = batch
        nq, sample = B, f'whole batch B={B}, median of 3 after 1 warm-up'
        if w['nppf'] >= 100:
            cb = {k: v[:1].clone() for k, v in batch.items()}
            nq, sample = 1, 'first query of the batch (B=1), median of 3 after 1 warm-up'
        cstep, ckind = _reference_step_fn(w, sd_cpu, 'cpu')             # cpu_baseline leg only
        ts = []
        for i in range(4):
            t0 = time.perf_counter()
            cstep(cb)
            if i:
                ts.append(time.perf_counter() - t0)
        line['cpu_baseline'] = {'value': nq / statistics.median(ts), 'unit': 'queries/s', 'cores': cores,
                                'kind': ckind, 'sample': sample}
    line.update(extras)
    line['ms_per_step_median'] = statistics.median(per_step)
    line['ms_per_step_min'] = min(per_step)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
