"""Four-modality MultiMAE (reference: pretraining/multimae/multimae_quadruplet.py): the plain zorro encoder
of multimae.py with a fourth input, the semantic class map `dnw` (SemSegInputAdapter, masked cross-entropy loss;
pretrain_mmae_my.py:45-84).  Token order s1, s2, dem, dnw, fusion (:397-417) -- a five-entry segment table for the same
kernels; token types as in zorro_utils_quadruplet.py.
forward(...) -> (preds, task_masks, return_tokens [B,R,D], ori_tokens [B,nenc,D], encoder_fusion_tokens [B,F,D])."""
from typing import Dict, Optional, Tuple

import torch.nn as nn

from ._core import MultiMAEBase
from .zorro_utils import LayerNorm
from .zorro_utils_quadruplet import TokenTypes


class MultiMAE(MultiMAEBase):
    FUSION_BLOCKS = False
    MODALITIES = ('s1', 's2', 'dem', 'dnw')
    TYPE_IDS = {'s1': TokenTypes.S1.value, 's2': TokenTypes.S2.value, 'dem': TokenTypes.DEM.value, 'dnw': TokenTypes.DNW.value}
    FUSION_TYPE_ID = TokenTypes.FUSION.value

    def __init__(self, input_adapters: Dict[str, nn.Module], output_adapters: Optional[Dict[str, nn.Module]],
                 num_global_tokens: int = 1, dim_tokens: int = 768, depth: int = 12, dim_head: int = 64, heads: int = 8,
                 ff_mult: int = 4, num_fusion_tokens: int = 16,
                 return_token_types: Tuple[TokenTypes] = (TokenTypes.S1, TokenTypes.S2, TokenTypes.DEM, TokenTypes.FUSION),
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = LayerNorm):
        super().__init__(input_adapters, output_adapters, num_global_tokens, dim_tokens, depth, dim_head, heads, ff_mult,
                         num_fusion_tokens, return_token_types, drop_path_rate, norm_layer)


def _factory(dim_tokens, depth, heads):
    def make(input_adapters: Dict[str, nn.Module], output_adapters: Optional[Dict[str, nn.Module]], **kwargs):
        return MultiMAE(input_adapters=input_adapters, output_adapters=output_adapters, dim_tokens=dim_tokens, depth=depth,
                        dim_head=64, heads=heads, ff_mult=4, norm_layer=LayerNorm, **kwargs)
    return make


# multimae_quadruplet.py:493-545
pretrain_multimae_tiny = _factory(384, 12, 8)
pretrain_multimae_base = _factory(768, 12, 8)
pretrain_multimae_large = _factory(1024, 24, 8)
