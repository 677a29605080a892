"""The decoder every reference script uses (reference: pretraining/multimae/output_adapters_simple.py):
proj_context -> + task embedding -> `depth` ViT blocks (dim 256) -> out_proj -> un-patchify.
Same constructor, `init`, `forward(encoder_tokens, input_info, ids_keep, ids_restore)` and state_dict
keys; Linear / LayerNorm / attention / un-patchify run on the sm_100a kernels."""
from functools import partial
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from .. import functions as Fn
from .multimae_utils import Block, ResidualStream, build_2d_sincos_posemb, pair, trunc_normal_


class SpatialOutputAdapter(nn.Module):
    def __init__(self, num_channels: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens_enc: Optional[int] = None, dim_tokens: int = 256, depth: int = 0,
                 learnable_pos_emb: int = False, image_size: Union[int, Tuple[int]] = 224, mlp_ratio: int = 4.0,
                 num_heads: int = 8, qkv_bias: bool = True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0,
                 drop_path_rate: float = 0.0, norm_layer: nn.Module = partial(nn.LayerNorm, eps=1e-6),
                 use_task_queries: bool = True, task: Optional[str] = None, context_tasks: Optional[list] = None,
                 use_xattn: bool = True):
        super().__init__()
        self.num_channels = num_channels
        self.stride_level = stride_level
        self.patch_size_full = pair(patch_size_full)
        self.dim_tokens_enc = dim_tokens_enc
        self.dim_tokens = dim_tokens
        self.learnable_pos_emb = learnable_pos_emb
        self.image_size = pair(image_size)
        self.use_task_queries = use_task_queries
        self.task = task
        self.use_xattn = use_xattn
        self.P_H = max(1, self.patch_size_full[0] // stride_level)
        self.P_W = max(1, self.patch_size_full[1] // stride_level)

        self.task_embeddings = None
        if context_tasks is not None:
            self.task_embeddings = nn.ParameterDict({t: nn.Parameter(torch.zeros(1, 1, self.dim_tokens)) for t in context_tasks})
            for emb in self.task_embeddings.values():
                trunc_normal_(emb, std=0.02)

        h = self.image_size[0] // (self.stride_level * self.P_H)
        w = self.image_size[1] // (self.stride_level * self.P_W)
        if not self.learnable_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h, w=w, embed_dim=self.dim_tokens), requires_grad=False)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, h, w, self.dim_tokens))
            trunc_normal_(self.pos_emb, std=0.02)

        if depth > 0:
            self.decoder_transformer = nn.Sequential(*[
                Block(dim=self.dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                      attn_drop=attn_drop_rate, drop_path=0.0 if drop_path_rate == 0 else drop_path_rate, norm_layer=norm_layer)
                for _ in range(depth)])
        else:
            self.decoder_transformer = nn.Identity()

        self.dim_patch = self.num_channels * self.P_H * self.P_W
        self.out_proj = nn.Linear(self.dim_tokens, self.dim_patch)
        if self.dim_tokens_enc is not None:
            self.init(dim_tokens_enc=dim_tokens_enc)

    def init(self, dim_tokens_enc: int = 768):
        self.dim_tokens_enc = dim_tokens_enc
        self.proj_context = nn.Linear(self.dim_tokens_enc, self.dim_tokens)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb', 'task_embeddings'}

    def forward(self, encoder_tokens: torch.Tensor, input_info: Dict, ids_keep: torch.Tensor = None,
                ids_restore: torch.Tensor = None):
        assert self.dim_tokens_enc is not None, 'Need to call init(dim_tokens_enc) function first'
        H, W = input_info['image_size']
        B, N, _ = encoder_tokens.shape
        d = self.dim_tokens
        bias = self.proj_context.bias
        if self.task_embeddings is not None and self.task in self.task_embeddings:
            bias = bias + self.task_embeddings[self.task].reshape(d)   # broadcast add folded into the GEMM bias
        x = Fn.linear(encoder_tokens.reshape(B * N, -1), self.proj_context.weight, bias, out_f32=True).view(B, N, d)
        x = self.decoder_transformer(ResidualStream.wrap(x)).tensor()
        x = Fn.linear(x.reshape(B * N, d), self.out_proj.weight, self.out_proj.bias)
        return Fn.UnpatchifyFn.apply(x, B, self.num_channels, H, W, self.P_H)
