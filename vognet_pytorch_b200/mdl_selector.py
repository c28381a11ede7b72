"""Drop-in for code/mdl_selector.py:26-69: ``get_mdl_loss_eval(cfg) -> {'mdl','loss','eval'}``
(classes, instantiated by the caller as ``mdl(cfg, comm)``, ``loss(cfg, comm)``,
``eval(cfg, comm, device)``: code/main_dist.py:33-53)."""
from . import mdl_vog
from .eval_vsrl_corr import EvaluatorSPAT, EvaluatorTEMP

_MODELS = {
    ('temp', 'igrnd'): mdl_vog.ImgGrnd_TEMP, ('temp', 'vgrnd'): mdl_vog.VidGrnd_TEMP,
    ('temp', 'vog'): mdl_vog.VOG_TEMP,
    ('spat', 'igrnd'): mdl_vog.ImgGrnd_SPAT, ('spat', 'vgrnd'): mdl_vog.VidGrnd_SPAT,
    ('spat', 'vog'): mdl_vog.VOG_SPAT,
}
_EVALS = {'temp': EvaluatorTEMP, 'spat': EvaluatorSPAT}


class _LossNotBuilt:
    """LossB_{TEMP,SPAT} (code/mdl_conc_single.py:180-433) is the first 'next' row (SURVEY.md
    section 8f): the forward-only build exports the name so the selector's dict shape is kept and
    fails loudly if a training loop tries to instantiate it."""

    def __init__(self, *a, **k):
        raise NotImplementedError('LossB_* is not built yet (forward/inference scope); '
                                  'SURVEY.md section 8f row 1')


def get_mdl_loss_eval(cfg):
    conc_type, mdl_type = cfg.ds.conc_type, cfg.mdl.name
    if conc_type in ('sep', 'svsq'):
        raise NotImplementedError("conc_type 'sep'/'svsq' (code/mdl_conc_sep.py) is outside the "
                                  'hot-path scope of this build (SURVEY.md section 2 row 7)')
    if (conc_type, mdl_type) not in _MODELS:
        raise NotImplementedError((conc_type, mdl_type))
    return {'mdl': _MODELS[(conc_type, mdl_type)], 'loss': _LossNotBuilt, 'eval': _EVALS[conc_type]}
