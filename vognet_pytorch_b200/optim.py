"""Flat-buffer Adam: the optimizer half of the reference's training step -
``torch.optim.Adam(params, lr=cfg.train.lr, betas=(0.9, 0.99))`` (code/main_dist.py:55, utils/trn_utils.py:799-803) -
as ONE kernel over all parameters and, for data-parallel training, ONE all-reduce over all gradients.

``FlatAdam(params, lr, betas, eps)`` moves every parameter into a single contiguous fp32 buffer (each ``param.data``
becomes a view of it, so the module keeps working and ``state_dict()`` is unchanged) and gives every parameter a
``.grad`` view of a second flat buffer.  ``step()`` launches ``vog_adam_step`` once; ``allreduce_grads()`` sums the flat
gradient across ranks with a single collective (NCCL over NVLink / NVSwitch on CUDA tensors) and folds the 1/world
into the following step.  The reference's DistributedDataParallel does the same reduction in ~25 MB buckets
(code/main_dist.py:76-85); parameters that received no gradient contribute zeros, which is what DDP's
``find_unused_parameters=True`` amounts to."""
import torch

from . import _lib, ops


class FlatAdam:
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), eps=1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('FlatAdam: no trainable parameters')
        dev = self.params[0].device
        if dev.type != 'cuda' or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise RuntimeError('FlatAdam: all parameters must be fp32 tensors on one CUDA device (no CPU path)')
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
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

    def zero_grad(self):
        self.flat_grad.zero_()           # the .grad views stay attached

    def allreduce_grads(self, group=None):
        """Sum the flat gradient over the ranks (one collective); the mean's 1/world is applied inside step()."""
        from . import runtime
        self._grad_scale = runtime.allreduce_flat_sum_(self.flat_grad, group)
        return self

    def step(self):
        self.step_count += 1
        L = _lib.lib()
        _lib.check(L.vog_adam_step(ops._ptr(self.flat_param), ops._ptr(self.flat_grad), ops._ptr(self.exp_avg),
                                   ops._ptr(self.exp_avg_sq), self.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                                   self.step_count, self._grad_scale, ops._stream()), 'vog_adam_step')
        self._grad_scale = 1.0

    def state_dict(self):
        return {'step': self.step_count, 'lr': self.lr, 'betas': self.betas, 'eps': self.eps,
                'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        self.step_count = int(sd['step'])
        self.lr, self.betas, self.eps = float(sd['lr']), tuple(sd['betas']), float(sd['eps'])
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
