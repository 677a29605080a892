"""MultiMAE with per-layer modality attention (`fus_blocks`, `mask_embedding`) and three per-modality
return tokens -- the model `pretrain_mmae.py:35` trains (reference: pretraining/multimae/
multimae_crossattn.py, with the working Block_Fusion of the downstream tree).
forward(...) returns the plain variant's 5-tuple + (return_token_s1, return_token_s2, return_token_dem), each [B,1,D]."""
from ._core import MultiMAEBase
from .multimae import _factory


class MultiMAE(MultiMAEBase):
    FUSION_BLOCKS = True


# multimae_crossattn.py:548-599: tiny is d=192 / 3 heads in this file
pretrain_multimae_tiny = _factory(MultiMAE, 192, 12, 3)
pretrain_multimae_base = _factory(MultiMAE, 768, 12, 8)
pretrain_multimae_large = _factory(MultiMAE, 1024, 24, 8)
