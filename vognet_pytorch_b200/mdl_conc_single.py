"""Loss-level boundary: ``LossB_SPAT`` / ``LossB_TEMP`` with the reference's
constructor ``(cfg, comm)`` and ``forward(out, inp) -> {'loss', 'mdl_out_loss'}`` (code/mdl_conc_single.py:
180-433), the arithmetic running in ``vog_loss_fwd``: IoU targets against the gt boxes of every SRL argument
(utils/box_utils.py:61-118), BCE with logits, masked mean * number of proposals * ``cfg.loss.loss_lambda``.

Forward only: the returned scalars carry no autograd graph (the backward of the path is a later row), which is
what validation (`code/eval_vsrl_corr.py:118-123`, `utils/trn_utils.py:443-483`) needs; training through it is
rejected loudly instead of silently producing zero gradients.
"""
import torch
from torch import nn

from . import ops


class _LossB(nn.Module):
    SPAT = True

    def __init__(self, cfg, comm):
        super().__init__()
        self.cfg, self.comm = cfg, comm
        self.loss_keys = ['loss', 'mdl_out_loss']
        self.loss_lambda = float(cfg.loss.loss_lambda)
        self.num_sampled_frm = int(cfg.ds.num_sampled_frm)
        self.num_prop_per_frm = int(comm['num_prop_per_frm'])

    def compute_loss_targets(self, inp, mdl_outs=None):
        """-> {'targets_one': bool [B,1,nsrl,P]} (code/mdl_conc_single.py:244-270,342-371)."""
        _, tg = self._run(inp, mdl_outs, want_targets=True)
        return {'targets_one': tg.unsqueeze(1)}

    def _run(self, inp, mdl_outs, want_targets=False):
        props = inp['pad_proposals']
        if not props.is_cuda:
            raise RuntimeError('vognet_pytorch_b200 runs on CUDA only (no CPU path)')
        B, P = props.shape[:2]
        nsrl = inp['srl_boxes'].shape[2]
        if inp['srl_boxes'].shape[1] != 1:
            raise NotImplementedError('temp/spat concatenation has one verb slot per query')
        if mdl_outs is None:
            mdl_outs = props.new_zeros(B, 1, nsrl, P)
        ncmp = inp['new_srl_idxs'].shape[1]
        res = ops.loss_fwd(mdl_outs.detach().reshape(B, nsrl, P).float(), props, inp['pad_gt_bboxs'],
                           inp['pad_frm_mask'], inp['pad_pnt_mask'], inp['srl_boxes'], inp['srl_boxes_lens'],
                           inp['srl_arg_boxes_mask'].reshape(B, nsrl), inp['num_cmp_msk'], inp['target_cmp'].reshape(B),
                           ncmp, self.num_prop_per_frm, self.SPAT, self.loss_lambda, want_targets=want_targets)
        return res if want_targets else (res, None)

    def forward(self, out, inp):
        mdl_outs = out['mdl_outs']
        if torch.is_grad_enabled() and mdl_outs.requires_grad:
            raise NotImplementedError('vognet_pytorch_b200: forward-only loss (no backward yet, SURVEY.md section 8f); '
                                      'call it under torch.no_grad()')
        loss, _ = self._run(inp, mdl_outs)
        loss = loss.reshape(())
        return {'loss': loss, 'mdl_out_loss': loss.clone()}


class LossB_SPAT(_LossB):
    """code/mdl_conc_single.py:335-433"""
    SPAT = True


class LossB_TEMP(_LossB):
    """code/mdl_conc_single.py:180-332"""
    SPAT = False
