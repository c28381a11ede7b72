"""Loss-level boundary: ``LossB_SPAT`` / ``LossB_TEMP`` with the reference's
constructor ``(cfg, comm)`` and ``forward(out, inp) -> {'loss', 'mdl_out_loss'}`` (code/mdl_conc_single.py:
180-433), the arithmetic running in ``vog_loss_fwd``: IoU targets against the gt boxes of every SRL argument
(utils/box_utils.py:61-118), BCE with logits, masked mean * number of proposals * ``cfg.loss.loss_lambda``.

The loss is differentiable with respect to the logits: when ``out['mdl_outs']`` requires grad the value comes from an
autograd Function whose backward is ``vog_loss_bwd`` (d loss / d logits, the first link of the backward chain of
SURVEY.md section 8f row 2; the model's own backward is not built, so in practice this serves callers that hold
logits as a leaf).  Under ``torch.no_grad()`` - validation, `code/eval_vsrl_corr.py:118-123`,
`utils/trn_utils.py:443-483` - nothing is saved.
"""
import torch
from torch import nn

from . import ops


class _GroundingLoss(torch.autograd.Function):
    """loss = vog_loss_fwd(logits, ...); backward = vog_loss_bwd on the targets / mask / statistics the forward left."""

    @staticmethod
    def forward(ctx, logits, args, kw):
        loss, tg, ws = ops.loss_fwd(logits, *args, want_targets=True, keep_ws=True, **kw)
        ctx.save_for_backward(logits, tg, ws)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        logits, tg, ws = ctx.saved_tensors
        return ops.loss_bwd(logits, tg, ws, grad_out.float()), None, None


def _differentiable(mdl_outs):
    return torch.is_grad_enabled() and mdl_outs is not None and mdl_outs.requires_grad


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
        args = (props, inp['pad_gt_bboxs'], inp['pad_frm_mask'], inp['pad_pnt_mask'], inp['srl_boxes'],
                inp['srl_boxes_lens'], inp['srl_arg_boxes_mask'].reshape(B, nsrl), inp['num_cmp_msk'],
                inp['target_cmp'].reshape(B), ncmp, self.num_prop_per_frm, self.SPAT)
        if _differentiable(mdl_outs) and not want_targets:
            return _GroundingLoss.apply(mdl_outs.reshape(B, nsrl, P).float(), args, dict(loss_lambda=self.loss_lambda)), None
        res = ops.loss_fwd(mdl_outs.detach().reshape(B, nsrl, P).float(), *args, self.loss_lambda, want_targets=want_targets)
        return res if want_targets else (res, None)

    def forward(self, out, inp):
        loss, _ = self._run(inp, out['mdl_outs'])
        loss = loss.reshape(())
        return {'loss': loss, 'mdl_out_loss': loss.clone()}


class LossB_SPAT(_LossB):
    """code/mdl_conc_single.py:335-433"""
    SPAT = True


class LossB_TEMP(_LossB):
    """code/mdl_conc_single.py:180-332"""
    SPAT = False


class LossB_SEP(nn.Module):
    """code/mdl_conc_sep.py:219-447: grounding loss of the SEP concatenation + the verb loss on the video-level
    logits.  ``forward(out, inp) -> {'loss', 'mdl_out_loss', 'verb_loss'}`` ('loss' is the grounding term only,
    :436-437).  Every (query, video) pair goes through ``vog_loss_fwd`` in its sep mode as one single-video problem
    whose overlaps survive only for the query's target video (:301-314); forward only, like LossB_SPAT/TEMP."""

    def __init__(self, cfg, comm):
        super().__init__()
        self.cfg, self.comm = cfg, comm
        self.loss_keys = ['loss', 'mdl_out_loss', 'verb_loss']
        self.loss_lambda = float(cfg.loss.loss_lambda)
        self.num_prop_per_frm = int(comm['num_prop_per_frm'])

    def _run(self, inp, mdl_outs, want_targets=False):
        props = inp['pad_proposals']
        if not props.is_cuda:
            raise RuntimeError('vognet_pytorch_b200 runs on CUDA only (no CPU path)')
        if props.dim() != 4:
            raise ValueError(f"conc_type 'sep' expects pad_proposals [B,ncmp,P1,7], got {tuple(props.shape)}")
        B, ncmp, P1 = props.shape[:3]
        Bq = B * ncmp
        sb, sl, am = inp['srl_boxes'], inp['srl_boxes_lens'], inp['srl_arg_boxes_mask']
        if sb.shape[1] == 1 and ncmp > 1:                  # one sentence slot for all videos (:246-247,336-337)
            sb, sl, am = (t.expand(B, ncmp, *t.shape[2:]) for t in (sb, sl, am))
        nsrl = sb.shape[2]
        if mdl_outs is None:
            mdl_outs = props.new_zeros(B, ncmp, nsrl, P1)
        if inp['pad_pnt_mask'].dim() != 3:
            raise AssertionError('pad_pnt_mask must be [B,ncmp,P1] (code/mdl_conc_sep.py:276)')
        # pair (b,c): target video index 0 iff c is the query's target, else 1 (never matches -> all-zero overlaps)
        tc = (torch.arange(ncmp, device=props.device).view(1, ncmp) != inp['target_cmp'].view(B, 1)).long().reshape(Bq)
        # the argument mask only decides masked-vs-plain mean (srl_arg_boxes_mask.max() > 0, :358-363)
        args = (props.reshape(Bq, P1, -1), inp['pad_gt_bboxs'].reshape(Bq, *inp['pad_gt_bboxs'].shape[2:]),
                inp['pad_frm_mask'].reshape(Bq, P1, -1), inp['pad_pnt_mask'].reshape(Bq, P1),
                sb.reshape(Bq, nsrl, -1), sl.reshape(Bq, nsrl, -1), am.reshape(Bq, nsrl),
                inp['num_cmp_msk'].reshape(Bq, 1), tc, 1, self.num_prop_per_frm, 2)
        if _differentiable(mdl_outs) and not want_targets:
            return _GroundingLoss.apply(mdl_outs.reshape(Bq, nsrl, P1).float(), args, dict(loss_lambda=self.loss_lambda)), None
        res = ops.loss_fwd(mdl_outs.detach().reshape(Bq, nsrl, P1).float(), *args, self.loss_lambda,
                           want_targets=want_targets)
        if want_targets:
            return res[0], res[1].view(B, ncmp, nsrl, P1)
        return res, None

    def compute_loss_targets(self, inp, mdl_outs=None):
        """-> {'targets_one': bool [B,ncmp,nsrl,P1]} (code/mdl_conc_sep.py:293-321)."""
        _, tg = self._run(inp, mdl_outs, want_targets=True)
        return {'targets_one': tg}

    def forward(self, out, inp):
        loss, _ = self._run(inp, out['mdl_outs'])
        loss = loss.reshape(())
        vidf = out['vidf_outs'].detach().float()         # verb_loss is reported only: 'loss' excludes it (:436-437)
        n = vidf.numel()
        verb = ops.verb_loss_fwd(vidf.reshape(n), inp['verb_cmp'].reshape(n),
                                 inp['verb_cross_cmp_msk'].reshape(n, -1), self.loss_lambda).reshape(())
        return {'loss': loss, 'mdl_out_loss': loss.clone(), 'verb_loss': verb}
