"""Plain zorro-encoder MultiMAE (reference: pretraining/multimae/multimae.py).
forward(...) -> (preds, task_masks, return_tokens [B,R,D], ori_tokens [B,nenc,D], encoder_fusion_tokens [B,F,D])
or (tokens, return_tokens, task_masks) when built without output adapters."""
from typing import Dict, Optional

import torch.nn as nn

from ._core import MultiMAEBase
from .zorro_utils import LayerNorm


class MultiMAE(MultiMAEBase):
    FUSION_BLOCKS = False


def _factory(cls, dim_tokens, depth, heads):
    def make(input_adapters: Dict[str, nn.Module], output_adapters: Optional[Dict[str, nn.Module]], **kwargs):
        return cls(input_adapters=input_adapters, output_adapters=output_adapters, dim_tokens=dim_tokens, depth=depth,
                   dim_head=64, heads=heads, ff_mult=4, norm_layer=LayerNorm, **kwargs)
    return make


# multimae.py:490-540: tiny is d=384 / 8 heads in this file
pretrain_multimae_tiny = _factory(MultiMAE, 384, 12, 8)
pretrain_multimae_base = _factory(MultiMAE, 768, 12, 8)
pretrain_multimae_large = _factory(MultiMAE, 1024, 24, 8)
