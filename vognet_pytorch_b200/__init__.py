"""vognet_pytorch_b200 - B200-native VOGNet forward fusion path (obj_tx + mul_tx with
relative-position bias) behind the reference's mdl_selector / nn.Module API.  See DESIGN.md."""
from . import synth  # noqa: F401
from .mdl_selector import get_mdl_loss_eval  # noqa: F401
from .mdl_vog import VOG_SEP, VOG_SPAT, VOG_TEMP  # noqa: F401
from .transformer_code import RelBias, RelTransformer, Transformer  # noqa: F401

__all__ = ['get_mdl_loss_eval', 'VOG_SPAT', 'VOG_TEMP', 'VOG_SEP', 'RelTransformer', 'Transformer', 'RelBias',
           'synth']
