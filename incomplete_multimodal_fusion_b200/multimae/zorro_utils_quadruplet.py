"""Encoder blocks of the 4-modality variant (reference: pretraining/multimae/zorro_utils_quadruplet.py): the same Block /
Attention / LayerNorm / Mlp as zorro_utils.py, with the token types renumbered for the semantic modality `dnw`
(:18-23: S1 0, S2 1, DEM 2, DNW 3, FUSION 4)."""
from enum import Enum

from .zorro_utils import Attention, Block, LayerNorm, Mlp, ZorroMask, block_params, exists  # noqa: F401


class TokenTypes(Enum):
    S1 = 0
    S2 = 1
    DEM = 2
    DNW = 3
    FUSION = 4
