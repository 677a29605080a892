"""Decoder-side building blocks with the reference's names (reference: pretraining/multimae/
multimae_utils.py).  Parameters keep the reference's names/shapes (state_dict compatible); the math
runs on the sm_100a kernels: LayerNorm -> fused kernel, Linear -> tcgen05 GEMM with bias / GELU /
residual epilogues, attention -> flash kernel (dh = dim / heads, scale applied to the scores)."""
import math
import warnings

import torch
import torch.nn as nn

from .. import functions as Fn


def pair(t):
    return t if isinstance(t, tuple) else (t, t)


def build_2d_sincos_posemb(h, w, embed_dim=1024, temperature=10000.):
    """[1, embed_dim, h, w] MoCo-v3 style table; layout quirks as multimae_utils.py:29-45 (w-major grid)."""
    assert embed_dim % 4 == 0, 'Embed dimension must be divisible by 4 for 2D sin-cos position embedding'
    ww, hh = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32), indexing='ij')
    quarter = embed_dim // 4
    freq = 1. / (temperature ** (torch.arange(quarter, dtype=torch.float32) / quarter))
    aw = ww.reshape(-1, 1) * freq.reshape(1, -1)
    ah = hh.reshape(-1, 1) * freq.reshape(1, -1)
    table = torch.cat([aw.sin(), aw.cos(), ah.sin(), ah.cos()], dim=1)
    return table.reshape(1, h, w, embed_dim).permute(0, 3, 1, 2).contiguous()


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """truncated normal init in place (multimae_utils.py:84-102)"""
    if (mean < a - 2 * std) or (mean > b + 2 * std):
        warnings.warn("mean is more than 2 std from [a, b] in trunc_normal_", stacklevel=2)
    with torch.no_grad():
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


def _flat(x):
    return x.reshape(-1, x.shape[-1])


class Mlp(nn.Module):
    """fc1 -> GELU -> fc2 with biases (multimae_utils.py:138-155).  2-D or 3-D input; bf16 output."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU or drop != 0.:
            raise NotImplementedError("only GELU / drop=0 are built (all the reference ever uses)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x, residual=None):
        shape = x.shape
        h = Fn.linear(_flat(x), self.fc1.weight, self.fc1.bias, act=1)
        y = Fn.linear(h, self.fc2.weight, self.fc2.bias, residual=None if residual is None else _flat(residual))
        return y.view(*shape[:-1], y.shape[-1])


class Attention(nn.Module):
    """ViT self-attention with a fused qkv projection (multimae_utils.py:158-182)"""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("dropout is never active in the reference path")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        if self.head_dim not in (32, 64):
            raise NotImplementedError("attention kernels are built for head_dim 32 and 64")
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, residual=None):
        """x: [B, N, C] bf16 (already normalised).  Returns proj(attn) (+ residual, then f32)."""
        B, N, C = x.shape
        qkv = Fn.linear(_flat(x), self.qkv.weight, self.qkv.bias)
        o = Fn.SelfAttentionFn.apply(qkv, B, N, self.num_heads, self.head_dim, self.scale)
        y = Fn.linear(o, self.proj.weight, self.proj.bias, residual=None if residual is None else _flat(residual))
        return y.view(B, N, C)


class CrossAttention(nn.Module):
    """queries attend to a context sequence (multimae_utils.py:185-214): q / kv / proj Linears, scale on the scores"""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0., proj_drop=0.):
        super().__init__()
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("dropout is never active in the reference path")
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        if self.head_dim not in (32, 64):
            raise NotImplementedError("attention kernels are built for head_dim 32 and 64")
        self.scale = self.head_dim ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, context):
        """x [B, N, C], context [B, M, C] (already normalised, bf16 or f32) -> proj(attention) bf16 [B, N, C]"""
        B, N, C = x.shape
        M = context.shape[1]
        q = Fn.linear(_flat(x), self.q.weight, self.q.bias)
        kv = Fn.linear(_flat(context), self.kv.weight, self.kv.bias)
        o = Fn.CrossAttentionFn.apply(q, kv, B, N, M, self.num_heads, self.head_dim, self.scale)
        return Fn.linear(o, self.proj.weight, self.proj.bias).view(B, N, C)


class ResidualStream:
    """fp32 residual stream [rows, C] plus the bf16 output of the last sub-layer that has not been added yet: the next
    LayerNorm launch adds it (functions.add_layer_norm), so no GEMM epilogue does an fp32 read-modify-write.
    `Block` accepts and returns this carrier when chained; given a plain tensor it returns a plain tensor."""

    def __init__(self, x, pend, shape):
        self.x, self.pend, self.shape = x, pend, tuple(shape)

    @staticmethod
    def wrap(t):
        return ResidualStream(_flat(t if t.dtype == torch.float32 else t.float()), None, t.shape)

    def tensor(self):
        """fp32 [..., C] with the pending delta added"""
        x = self.x if self.pend is None else Fn.AddDeltaFn.apply(self.x, self.pend)
        return x.view(self.shape)


class Block(nn.Module):
    """pre-LN ViT block of the decoders (multimae_utils.py:217-232); residual stream in fp32"""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if drop_path != 0.:
            raise NotImplementedError("stochastic depth is not built (rate 0 in every reference script)")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        chained = isinstance(x, ResidualStream)
        st = x if chained else ResidualStream.wrap(x)
        B, N, C = st.shape
        if st.pend is None:
            x32 = st.x
            h = Fn.layer_norm(x32, self.norm1.weight, self.norm1.bias, self.norm1.eps, out_bf16=True)
        else:   # x + (previous sub-layer) and norm1 in one pass
            x32, h = Fn.add_layer_norm(st.x, st.pend, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        d_attn = self.attn(h.view(B, N, C))                                   # bf16, no residual in the GEMM
        x32, h = Fn.add_layer_norm(x32, _flat(d_attn), self.norm2.weight, self.norm2.bias, self.norm2.eps)
        out = ResidualStream(x32, _flat(self.mlp(h.view(B, N, C))), st.shape)
        return out if chained else out.tensor()
