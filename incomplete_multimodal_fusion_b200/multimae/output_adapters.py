"""The cross-attention decoder the reference package exports as ``multimae.SpatialOutputAdapter``
(reference: pretraining/multimae/output_adapters.py; ``output_adapters_fusion.py`` is the same class).

  proj_context -> pad with mask tokens and un-shuffle (ids_restore) -> + task / sin-cos position embeddings
  -> queries = this task's slice, context = the visible tokens again (ids_keep)
  -> CrossAttention(query_norm(q), context_norm(ctx)) -> x + Mlp(out_norm(x)) -> `depth` ViT blocks -> out_proj
  -> un-patchify.

Same constructor, `init`, `forward(encoder_tokens, input_info, ids_keep, ids_restore)`, helper methods and
state_dict keys as the reference.  Linear / LayerNorm / attention / un-patchify run on the sm_100a kernels; the
two token gathers (int64 index plumbing on [B, n, 256] tensors) stay torch.gather."""
from functools import partial
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functions as Fn
from .multimae_utils import Block, CrossAttention, Mlp, ResidualStream, build_2d_sincos_posemb, pair, trunc_normal_


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
        self.mask_token = nn.Parameter(torch.zeros(1, 1, self.dim_tokens))

        h = self.image_size[0] // (self.stride_level * self.P_H)
        w = self.image_size[1] // (self.stride_level * self.P_W)
        if not self.learnable_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h, w=w, embed_dim=self.dim_tokens), requires_grad=False)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, h, w, self.dim_tokens))
            trunc_normal_(self.pos_emb, std=0.02)

        if self.use_xattn:
            self.decoder = CrossAttention(dim=self.dim_tokens, num_heads=num_heads, qkv_bias=qkv_bias,
                                          attn_drop=attn_drop_rate, proj_drop=drop_rate)
            self.context_norm = norm_layer(self.dim_tokens)
            self.query_norm = norm_layer(self.dim_tokens)
            self.out_norm = norm_layer(self.dim_tokens)
            self.mlp = Mlp(in_features=self.dim_tokens, hidden_features=int(self.dim_tokens * mlp_ratio))

        if depth > 0:
            if drop_path_rate != 0:
                raise NotImplementedError("stochastic depth is not built (rate 0 in every reference script)")
            self.decoder_transformer = nn.Sequential(*[
                Block(dim=self.dim_tokens, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                      attn_drop=attn_drop_rate, drop_path=0.0, norm_layer=norm_layer) for _ in range(depth)])
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
        return {'pos_emb', 'mask_token', 'task_embeddings'}

    def _pos_tokens(self, size):
        """[1, nh*nw, d] position table (bilinear resize is the identity at the native grid, Appendix A #9)"""
        pe = self.pos_emb
        if tuple(pe.shape[-2:]) != tuple(size):
            pe = F.interpolate(pe, size=size, mode='bilinear', align_corners=False)
        return pe.flatten(2).transpose(1, 2)

    def generate_context_embeddings(self, input_info, bs: int, size: Tuple[int, int], device: Optional[torch.device] = None):
        """task embedding + position embedding of every token slot, in input_info order (output_adapters.py:160-181)"""
        out = []
        for task, info in input_info['tasks'].items():
            if self.task_embeddings is not None and task in self.task_embeddings:
                emb = self.task_embeddings[task].expand(bs, info['num_tokens'], self.dim_tokens)
            else:
                emb = torch.zeros((bs, info['num_tokens'], self.dim_tokens), device=device)
            if info['has_2d_posemb']:
                pos = self._pos_tokens(size)
                assert info['num_tokens'] == pos.shape[1]
                emb = emb + pos
            out.append(emb)
        return torch.cat(out, dim=1)

    def get_queries_and_context(self, context_tokens, input_info, ids_keep, ids_restore):
        """output_adapters.py:183-234"""
        B, _, d = context_tokens.shape
        H, W = input_info['image_size']
        N_H = H // (self.stride_level * self.P_H)
        N_W = W // (self.stride_level * self.P_W)
        n_glob = input_info.get('num_global_tokens', 0)
        ctx = context_tokens[:, :-n_glob] if n_glob else context_tokens
        if ids_restore.shape[0] != B:
            ids_restore = ids_restore.expand(B, -1)
        if ids_keep.shape[0] != B:
            ids_keep = ids_keep.expand(B, -1)
        mask_tokens = self.mask_token.to(ctx.dtype).expand(B, input_info['num_task_tokens'] - ctx.shape[1], d)
        full = torch.cat([ctx, mask_tokens], dim=1)
        full = torch.gather(full, 1, ids_restore.unsqueeze(-1).expand(-1, -1, d))
        full = full + self.generate_context_embeddings(input_info=input_info, bs=B, size=(N_H, N_W), device=ctx.device)
        if self.use_task_queries and self.task in input_info['tasks']:
            info = input_info['tasks'][self.task]
            queries = full[:, info['start_idx']:info['end_idx']]
        else:
            queries = self.mask_token.expand(B, N_H * N_W, d) + self._pos_tokens((N_H, N_W))
            if self.task_embeddings is not None and self.task in self.task_embeddings:
                queries = queries + self.task_embeddings[self.task]
        ctx = torch.gather(full, 1, ids_keep.unsqueeze(-1).expand(-1, -1, d))
        if n_glob:
            ctx = torch.cat([ctx, context_tokens[:, -n_glob:]], dim=1)
        return queries, ctx

    def forward(self, encoder_tokens: torch.Tensor, input_info: Dict, ids_keep: torch.Tensor, ids_restore: torch.Tensor):
        assert self.dim_tokens_enc is not None, 'Need to call init(dim_tokens_enc) function first'
        H, W = input_info['image_size']
        B, N, _ = encoder_tokens.shape
        d = self.dim_tokens
        ctx = Fn.linear(encoder_tokens.reshape(B * N, -1), self.proj_context.weight, self.proj_context.bias, out_f32=True)
        queries, ctx = self.get_queries_and_context(ctx.view(B, N, d), input_info, ids_keep, ids_restore)
        Nq, Nk = queries.shape[1], ctx.shape[1]
        if self.use_xattn:
            qn = Fn.layer_norm(queries.reshape(B * Nq, d), self.query_norm.weight, self.query_norm.bias, self.query_norm.eps,
                               out_bf16=True).view(B, Nq, d)
            cn = Fn.layer_norm(ctx.reshape(B * Nk, d), self.context_norm.weight, self.context_norm.bias,
                               self.context_norm.eps, out_bf16=True).view(B, Nk, d)
            x = self.decoder(qn, cn).float()                        # no residual around the cross-attention (:265)
            h = Fn.layer_norm(x.reshape(B * Nq, d), self.out_norm.weight, self.out_norm.bias, self.out_norm.eps,
                              out_bf16=True).view(B, Nq, d)
            st = ResidualStream(x.reshape(B * Nq, d), self.mlp(h).reshape(B * Nq, d), (B, Nq, d))   # x + Mlp(out_norm(x)) (:266)
        else:
            st = ResidualStream.wrap(queries)
        x = self.decoder_transformer(st).tensor()
        x = Fn.linear(x.reshape(B * Nq, d), self.out_proj.weight, self.out_proj.bias)
        return Fn.UnpatchifyFn.apply(x, B, self.num_channels, H, W, self.P_H)
