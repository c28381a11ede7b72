"""Flat-buffer Adam: the optimizer half of the reference's training step -
``torch.optim.Adam(params, lr=cfg.train.lr, betas=(0.9, 0.99))`` (code/main_dist.py:55, utils/trn_utils.py:799-803) -
as ONE kernel over all parameters and, for data-parallel training, ONE all-reduce over all gradients.

``FlatAdam(params, lr, betas, eps)`` is a ``torch.optim.Optimizer`` (single param group; ``LambdaLR`` /
``ReduceLROnPlateau`` of utils/trn_utils.py:807-818 wrap it and their lr changes are read every step).  It moves every
parameter into a single contiguous fp32 buffer (each ``param.data`` becomes a view of it, so the module keeps working
and ``state_dict()`` is unchanged) and gives every parameter a ``.grad`` view of a second flat buffer.  ``step()``
launches ``vog_adam_step`` once; ``allreduce_grads()`` sums the flat gradient across ranks with a single collective
(NCCL over NVLink / NVSwitch on CUDA tensors) and folds the 1/world into the following step.  The reference's
DistributedDataParallel does the same reduction in ~25 MB buckets (code/main_dist.py:76-85); parameters that received
no gradient contribute zeros, which is what DDP's ``find_unused_parameters=True`` amounts to.

``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's layout (per-parameter ``step`` / ``exp_avg`` /
``exp_avg_sq`` under integer ids + ``param_groups``), so the ``optimizer_state_dict`` of a reference checkpoint
(utils/trn_utils.py:610,622) resumes here and vice versa."""
import torch

from . import _lib, ops, packing


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), eps=1e-8):
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError('FlatAdam: no trainable parameters')
        dev = plist[0].device
        if dev.type != 'cuda' or any(p.device != dev or p.dtype != torch.float32 for p in plist):
            raise RuntimeError('FlatAdam: all parameters must be fp32 tensors on one CUDA device (no CPU path)')
        super().__init__(plist, dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps),
                                     weight_decay=0, amsgrad=False, maximize=False, foreach=None, capturable=False,
                                     differentiable=False, fused=None))
        self.params = plist
        self.step_count = 0
        self._grad_scale = 1.0
        # every tensor starts on a 16-byte boundary of the flat buffers (float4 accesses in the kernel)
        offs, n = [], 0
        for p in self.params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.flat_param = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.offsets = offs

    # the scalar hyper-parameters live in the (single) param group, where lr schedulers write them
    @property
    def lr(self):
        return float(self.param_groups[0]['lr'])

    @lr.setter
    def lr(self, v):
        self.param_groups[0]['lr'] = float(v)

    @property
    def betas(self):
        return tuple(float(b) for b in self.param_groups[0]['betas'])

    @property
    def eps(self):
        return float(self.param_groups[0]['eps'])

    def zero_grad(self, set_to_none=False):
        """Zero the flat gradient; the ``.grad`` views stay attached (set_to_none would detach them, so it is ignored)."""
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):           # re-attach views a caller may have dropped / replaced
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def allreduce_grads(self, group=None, async_op=False):
        """Sum the flat gradient over the ranks (one collective); the mean's 1/world is applied inside step()."""
        from . import runtime
        self._grad_scale, work = runtime.allreduce_flat_sum_(self.flat_grad, group, async_op=async_op)
        return work if async_op else self

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.step_count += 1
        L = _lib.lib()
        b1, b2 = self.betas
        _lib.check(L.vog_adam_step(ops._ptr(self.flat_param), ops._ptr(self.flat_grad), ops._ptr(self.exp_avg),
                                   ops._ptr(self.exp_avg_sq), self.numel, self.lr, b1, b2, self.eps,
                                   self.step_count, self._grad_scale, ops._stream()), 'vog_adam_step')
        self._grad_scale = 1.0
        # the kernel wrote the parameters through a raw pointer: neither data_ptr nor _version moved, so every
        # packed low-precision weight copy (and captured graph) has to be told
        packing.bump_generation()
        return loss

    # ---- torch.optim.Adam checkpoint layout ----------------------------------------------------
    def state_dict(self):
        st = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            st[i] = {'step': torch.tensor(float(self.step_count)),
                     'exp_avg': self.exp_avg[o:o + n].view_as(p).clone(),
                     'exp_avg_sq': self.exp_avg_sq[o:o + n].view_as(p).clone()}
        groups = []
        for g in self.param_groups:
            gg = {k: v for k, v in g.items() if k != 'params'}
            gg['params'] = list(range(len(self.params)))
            groups.append(gg)
        return {'state': st, 'param_groups': groups}

    def load_state_dict(self, sd):
        if 'state' not in sd:                                  # round-1 private layout
            self.step_count = int(sd['step'])
            self.param_groups[0].update(lr=float(sd['lr']), betas=tuple(sd['betas']), eps=float(sd['eps']))
            self.exp_avg.copy_(sd['exp_avg'])
            self.exp_avg_sq.copy_(sd['exp_avg_sq'])
            return
        ids = list(sd['param_groups'][0]['params']) if sd.get('param_groups') else list(range(len(self.params)))
        if len(ids) != len(self.params):
            raise ValueError(f'FlatAdam.load_state_dict: checkpoint has {len(ids)} parameters, optimizer has '
                             f'{len(self.params)}')
        steps = set()
        with torch.no_grad():
            for pid, p, o in zip(ids, self.params, self.offsets):
                ent = sd['state'].get(pid, sd['state'].get(str(pid)))
                n = p.numel()
                if ent is None:                                # parameter that never received a gradient
                    self.exp_avg[o:o + n].zero_()
                    self.exp_avg_sq[o:o + n].zero_()
                    continue
                if tuple(ent['exp_avg'].shape) != tuple(p.shape):
                    raise ValueError(f'FlatAdam.load_state_dict: state {pid} has shape {tuple(ent["exp_avg"].shape)}, '
                                     f'parameter has {tuple(p.shape)}')
                self.exp_avg[o:o + n].view_as(p).copy_(ent['exp_avg'])
                self.exp_avg_sq[o:o + n].view_as(p).copy_(ent['exp_avg_sq'])
                steps.add(int(float(ent['step'])))
        if len(steps) > 1:
            raise ValueError(f'FlatAdam.load_state_dict: parameters are at different steps {sorted(steps)} - one flat '
                             'buffer takes one step count')
        self.step_count = steps.pop() if steps else 0
        if sd.get('param_groups'):
            g = sd['param_groups'][0]
            for k in ('lr', 'betas', 'eps', 'initial_lr'):
                if k in g:
                    self.param_groups[0][k] = tuple(g[k]) if k == 'betas' else g[k]
