"""Two-modality (s2, dem) MultiMAE whose fusion tokens are one per VISIBLE token, initialised by a BiLSTM over the
(modality token, fusion token) pair -- the model `pretrain_mmae_s2dsm.py:35` trains (reference:
pretraining/multimae/multimae_lstm_s2dsm.py; BASELINE config 1).
forward(...) -> (preds, task_masks, return_tokens [B,3,D], ori_tokens [B,nenc,D], encoder_fusion_tokens [B,nenc,D]);
the decoders receive the full fusion grid [B,F,D] with the encoded fusion tokens scattered back."""
from typing import Tuple

from ._core import MultiMAEBase
from .multimae import _factory
from .zorro_utils import LayerNorm, TokenTypes


class MultiMAE(MultiMAEBase):
    FUSION_BLOCKS = False
    LSTM_FUSION = True
    MODALITIES = ('s2', 'dem')

    def __init__(self, input_adapters, output_adapters, num_global_tokens: int = 1, dim_tokens: int = 768, depth: int = 12,
                 dim_head: int = 64, heads: int = 8, ff_mult: int = 4, num_fusion_tokens: int = 16,
                 return_token_types: Tuple[TokenTypes] = (TokenTypes.S1, TokenTypes.S2, TokenTypes.DEM, TokenTypes.FUSION),
                 drop_path_rate: float = 0.0, norm_layer=LayerNorm):
        super().__init__(input_adapters, output_adapters, num_global_tokens=num_global_tokens, dim_tokens=dim_tokens,
                         depth=depth, dim_head=dim_head, heads=heads, ff_mult=ff_mult, num_fusion_tokens=num_fusion_tokens,
                         return_token_types=return_token_types, drop_path_rate=drop_path_rate, norm_layer=norm_layer)


# multimae_lstm_s2dsm.py:505-556: tiny is d=192 / 3 heads in this file
pretrain_multimae_tiny = _factory(MultiMAE, 192, 12, 3)
pretrain_multimae_base = _factory(MultiMAE, 768, 12, 8)
pretrain_multimae_large = _factory(MultiMAE, 1024, 24, 8)
