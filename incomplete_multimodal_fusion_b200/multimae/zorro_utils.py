"""Encoder primitives under the reference's names (reference: pretraining/multimae/zorro_utils.py,
working Block_Fusion from downstream/instance_segmentation/modeling/multimae/zorro_utils.py:243-258).

These modules own the parameters (same names / shapes as the reference, so checkpoints load with
strict=True).  `MultiMAE.forward` does not call them layer by layer: it hands all their parameters to
one fused autograd node (functions.EncoderStackFn).  Their own `forward`s route through the same node
so that they can be used standalone with `ZorroMask` (or no mask)."""
from enum import Enum

import torch
import torch.nn as nn

from .. import functions as Fn
from .multimae_utils import Mlp  # noqa: F401  (re-exported like the reference does)


class TokenTypes(Enum):
    S1 = 0
    S2 = 1
    DEM = 2
    FUSION = 3


def exists(val):
    return val is not None


def default(*args):
    for a in args:
        if exists(a):
            return a
    return None


class ZorroMask:
    """Structured form of the zorro attention mask (multimae.py:410-426): `counts` visible tokens per
    modality followed by `n_fusion` fusion tokens.  Same-type tokens attend to each other, fusion tokens
    attend to everything.  The kernels consume the segment table, never a dense [N, N] mask."""

    def __init__(self, counts, n_fusion, device):
        self.counts = [int(c) for c in counts]
        self.n_fusion = int(n_fusion)
        bounds = [0]
        for c in self.counts:
            bounds.append(bounds[-1] + c)
        bounds.append(bounds[-1] + self.n_fusion)
        self.bounds = bounds
        self.seg = torch.tensor(bounds, dtype=torch.int32, device=device)

    @property
    def nseg(self):
        return len(self.bounds) - 1

    def dense(self):
        types = torch.repeat_interleave(torch.arange(self.nseg, device=self.seg.device),
                                        torch.tensor(self.counts + [self.n_fusion], device=self.seg.device))
        return (types[:, None] == types[None, :]) | (types[:, None] == self.nseg - 1)


class LayerNorm(nn.Module):
    """bias-less LayerNorm: learnable gamma, zero `beta` buffer, eps 1e-5 (zorro_utils.py:103-110)"""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))
        self.register_buffer("beta", torch.zeros(dim))

    def forward(self, x):
        y = Fn.layer_norm(x.reshape(-1, x.shape[-1]).float(), self.gamma, None, 1e-5, out_bf16=False)
        return y.view(x.shape)


class GEGLU(nn.Module):
    """gelu(gate) * value on the two halves of the last dim (zorro_utils.py:115-118).  Inside the fused
    path this is a GEMM epilogue; the module exists for the state_dict index layout of FeedForward."""

    def forward(self, x):
        raise RuntimeError("GEGLU is fused into the FFN GEMM epilogue; call the FeedForward container")


class _FeedForward(nn.Sequential):
    """Sequential(LayerNorm, Linear(d, 2I), GEGLU, Linear(I, d)) -- keys 0.gamma, 0.beta, 1.weight, 3.weight"""

    def forward(self, x):
        raise RuntimeError("FeedForward runs inside functions.EncoderStackFn; use Block / Block_Fusion")


def FeedForward(dim, mult=4):
    inner = int(dim * mult * 2 / 3)
    return _FeedForward(LayerNorm(dim), nn.Linear(dim, inner * 2, bias=False), GEGLU(), nn.Linear(inner, dim, bias=False))


class Attention(nn.Module):
    """to_q / to_kv / to_out without biases, LayerNorm on the query side (zorro_utils.py:152-194)"""

    def __init__(self, dim, dim_head=64, heads=8):
        super().__init__()
        if dim_head != 64:
            raise NotImplementedError("encoder attention kernels are built for dim_head 64 (all reference factories)")
        self.scale = dim_head ** -0.5
        self.heads = heads
        inner = dim_head * heads
        self.norm = LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, context=None, attn_mask=None):
        raise RuntimeError("zorro Attention runs inside Block (self-attention) or the MultiMAE pooling head")


class Attention_LSTM(nn.Module):
    """softmax over time of Linear(tanh(H)) (zorro_utils.py:261-273)"""

    def __init__(self, hidden_size):
        super().__init__()
        self.attention = nn.Linear(hidden_size, 1)

    def forward(self, H, mask=None):
        M = self.attention(torch.tanh(H)).squeeze(2)
        if mask is not None:
            M = M.masked_fill(mask == 0, -1e+4)
        return torch.softmax(M, dim=1).unsqueeze(1)


class AttentionBiLSTM(nn.Module):
    """Bidirectional LSTM over a short sequence, directions summed, attention-pooled over time
    (zorro_utils.py:276-299).  Used by the BiLSTM-fusion variant on length-2 (token, fusion token) sequences; the
    recurrence stays the library LSTM (cuDNN) -- SURVEY.md 8a row a21 keeps it out of the custom-kernel scope."""

    def __init__(self, embedding_dim, num_layers=1, dropout=0.0, emb_layer_dropout=0.0):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.lstm = nn.LSTM(embedding_dim, embedding_dim, num_layers, dropout=(0 if num_layers == 1 else dropout),
                            bidirectional=True, batch_first=True)
        self.attention = Attention_LSTM(embedding_dim)

    def forward(self, embedded, mask=None):
        y, _ = self.lstm(embedded)
        y = y[:, :, :self.embedding_dim] + y[:, :, self.embedding_dim:]
        alpha = self.attention(y, mask)
        return alpha.bmm(y).squeeze(1)


def block_params(blk):
    """the 9 tensors EncoderStackFn expects for one Block / Block_Fusion, in its order"""
    return [blk.norm1.gamma, blk.attn.norm.gamma, blk.attn.to_q.weight, blk.attn.to_kv.weight, blk.attn.to_out.weight,
            blk.norm2.gamma, blk.mlp[0].gamma, blk.mlp[1].weight, blk.mlp[3].weight]


class Block(nn.Module):
    """x + Attn(norm1 x), x + FFN(norm2 x) with the double LayerNorms (zorro_utils.py:227-240)"""

    def __init__(self, dim=768, dim_head=64, heads=8, ff_mult=4, drop_path=0., norm_layer=LayerNorm):
        super().__init__()
        if drop_path > 0.:
            raise NotImplementedError("stochastic depth is not built (rate 0 in every reference script)")
        self.dim, self.heads, self.ff_inner = dim, heads, int(dim * ff_mult * 2 / 3)
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim=dim, dim_head=dim_head, heads=heads)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = FeedForward(dim=dim, mult=ff_mult)

    def forward(self, x, attn_mask=None):
        """x: [B, N, D] fp32 in the reference's concatenated order; attn_mask: None or ZorroMask."""
        B, N, D = x.shape
        if attn_mask is None:
            zm = ZorroMask([], N, x.device)
        elif isinstance(attn_mask, ZorroMask):
            zm = attn_mask
        else:
            raise NotImplementedError("pass a ZorroMask (segment structure); dense [N, N] masks are not consumed")
        nf = zm.n_fusion
        nenc = N - nf
        planar = torch.cat([x[:, :nenc].reshape(-1, D), x[:, nenc:].reshape(-1, D)], 0).float()
        meta = dict(B=B, D=D, H=self.heads, F=nf, nenc=nenc, fusion=False, depth=1, I=self.ff_inner, seg=zm.seg,
                    nseg=zm.nseg, slotmap=None)
        out = Fn.EncoderStackFn.apply(meta, planar, *block_params(self))
        return torch.cat([out[:B * nenc].view(B, nenc, D), out[B * nenc:].view(B, nf, D)], 1)


class Block_Fusion(nn.Module):
    """modality attention over the per-position slots, keep the fusion slot, then the FFN
    (downstream/.../zorro_utils.py:243-258).  Runs fused inside MultiMAE (functions.EncoderStackFn)."""

    def __init__(self, dim=768, dim_head=64, heads=8, ff_mult=4, norm_layer=LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.norm2 = norm_layer(dim)
        self.attn = Attention(dim=dim, dim_head=dim_head, heads=heads)
        self.mlp = FeedForward(dim=dim, mult=ff_mult)

    def forward(self, x, attn_mask=None):
        raise RuntimeError("Block_Fusion runs fused inside multimae_crossattn.MultiMAE (slot attention kernel)")
