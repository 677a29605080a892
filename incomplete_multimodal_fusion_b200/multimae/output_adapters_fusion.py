"""``output_adapters_fusion.py`` of the reference is the cross-attention decoder of ``output_adapters.py`` with one
difference: `no_weight_decay` (reference output_adapters_fusion.py:158).  Re-exported here under the same module name."""
import torch

from .output_adapters import SpatialOutputAdapter as _Base


class SpatialOutputAdapter(_Base):
    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb', 'task_embeddings'}
