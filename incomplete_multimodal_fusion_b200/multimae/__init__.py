"""Drop-in surface of the reference's ``multimae`` package (pretraining/multimae/__init__.py:1-4 exports
MaskedL1Loss, MaskedMSELoss, PatchedInputAdapter, MultiMAE, SpatialOutputAdapter).  Submodules keep the
reference's module names so ``from multimae.multimae_crossattn import pretrain_multimae_base`` style imports
port by changing the package prefix."""
from .criterion import MaskedL1Loss, MaskedMSELoss  # noqa: F401
from .input_adapters import FusionInputAdapter, PatchedInputAdapter  # noqa: F401
from .multimae import MultiMAE  # noqa: F401
from .output_adapters import SpatialOutputAdapter  # noqa: F401  (the cross-attention decoder, as in the reference's __init__)
