"""Selection step of the reference evaluators (code/eval_vsrl_corr.py:289-345 TEMP, :357-424 SPAT):
``get_out_results_boxes(out, inp) -> {'boxes','scores','indexs'}`` on the GPU through
``vog_select_fwd`` - per-(srl, frame, vid) max/argmax over the proposals of a frame, gather of the
winning 7-float proposal rows, argmax over the concatenated videos.  Bit-exact against torch.max /
gather / argmax given identical scores (lowest index wins ties).

Everything else in the reference evaluator (pickling predictions, cross-rank merge through files,
pandas accuracy computation: :101-150, code/eval_fn_corr.py) is file-I/O glue outside the hot path
 and is not rebuilt here; ``Evaluator*`` keep the constructor
signature ``(cfg, comm, device)`` so ``get_mdl_loss_eval`` stays source-compatible.
"""
from torch import nn

from . import ops


class _EvaluatorBase(nn.Module):
    SPAT = True

    def __init__(self, cfg, comm, device=None):
        super().__init__()
        self.cfg, self.comm, self.device = cfg, comm, device
        self.num_sampled_frm = int(cfg.ds.num_sampled_frm)
        self.num_frms = self.num_sampled_frm
        self.num_prop_per_frm = int(comm['num_prop_per_frm'])
        self.met_keys = ['avg1', 'avg1_cons', 'avg1_vidf', 'avg1_strict']

    def get_out_results_boxes(self, out_result_dict, inp):
        assert isinstance(out_result_dict, dict)
        scores = out_result_dict['mdl_outs_eval']
        B, nv, nsrl, P = scores.shape
        assert nv == 1
        ncmp = inp['new_srl_idxs'].size(1)
        boxes, sc, ix = ops.select_fwd(scores.reshape(B, nsrl, P), inp['pad_proposals'], ncmp,
                                       self.num_sampled_frm, self.num_prop_per_frm, self.SPAT)
        return {'boxes': boxes, 'scores': sc, 'indexs': ix}

    def forward(self, *a, **k):
        raise NotImplementedError('dataset-driven validation loop (code/eval_vsrl_corr.py:101-150) is '
                                  'outside the hot-path scope; use get_out_results_boxes')


class EvaluatorSPAT(_EvaluatorBase):
    SPAT = True


class EvaluatorTEMP(_EvaluatorBase):
    SPAT = False


class EvaluatorSEP(_EvaluatorBase):
    """code/eval_vsrl_corr.py:153-220: SEP outputs keep the [ncmp] axis; the predicted video is the argmax of the
    fused per-video score ``fin_scores``."""

    def get_out_results_boxes(self, out_result_dict, inp):
        assert isinstance(out_result_dict, dict)
        scores, fin = out_result_dict['mdl_outs_eval'], out_result_dict['fin_scores']
        boxes, sc, ix = ops.select_sep_fwd(scores, inp['pad_proposals'], fin, self.num_sampled_frm,
                                           self.num_prop_per_frm)
        return {'boxes': boxes, 'scores': sc, 'indexs': ix}
