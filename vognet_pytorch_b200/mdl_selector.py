"""Drop-in for code/mdl_selector.py:26-69: ``get_mdl_loss_eval(cfg) -> {'mdl','loss','eval'}``
(classes, instantiated by the caller as ``mdl(cfg, comm)``, ``loss(cfg, comm)``,
``eval(cfg, comm, device)``: code/main_dist.py:33-53)."""
from . import mdl_vog
from .eval_vsrl_corr import EvaluatorSEP, EvaluatorSPAT, EvaluatorTEMP
from .mdl_conc_single import LossB_SEP, LossB_SPAT, LossB_TEMP

_MODELS = {
    ('temp', 'igrnd'): mdl_vog.ImgGrnd_TEMP, ('temp', 'vgrnd'): mdl_vog.VidGrnd_TEMP,
    ('temp', 'vog'): mdl_vog.VOG_TEMP,
    ('spat', 'igrnd'): mdl_vog.ImgGrnd_SPAT, ('spat', 'vgrnd'): mdl_vog.VidGrnd_SPAT,
    ('spat', 'vog'): mdl_vog.VOG_SPAT,
    ('sep', 'igrnd'): mdl_vog.ImgGrnd_SEP, ('sep', 'vgrnd'): mdl_vog.VidGrnd_SEP,
    ('sep', 'vog'): mdl_vog.VOG_SEP,
}
_EVALS = {'temp': EvaluatorTEMP, 'spat': EvaluatorSPAT, 'sep': EvaluatorSEP}


_LOSSES = {'temp': LossB_TEMP, 'spat': LossB_SPAT, 'sep': LossB_SEP}


def get_mdl_loss_eval(cfg):
    conc_type, mdl_type = cfg.ds.conc_type, cfg.mdl.name
    if conc_type == 'svsq':          # same classes, one video per query (code/mdl_selector.py:29)
        conc_type = 'sep'
    if (conc_type, mdl_type) not in _MODELS:
        raise NotImplementedError((conc_type, mdl_type))
    return {'mdl': _MODELS[(conc_type, mdl_type)], 'loss': _LOSSES[conc_type],
            'eval': _EVALS[conc_type]}
