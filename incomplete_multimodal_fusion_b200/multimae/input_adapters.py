"""Input adapters with the reference's constructor / attribute / state_dict surface
(reference: pretraining/multimae/input_adapters.py).

`PatchedInputAdapter` owns the patch-projection Conv2d parameters and the sin-cos table.  Inside
`MultiMAE.forward` the projection is NOT run as a convolution over every patch: only the visible
patches are gathered (im2col) and projected by the tcgen05 GEMM with bias + pos-emb fused in the
epilogue (functions.EmbedFn).  Calling the adapter on its own embeds all patches through the same path.
"""
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from .. import functions as Fn
from .multimae_utils import build_2d_sincos_posemb, pair, trunc_normal_


class _SpatialAdapterBase(nn.Module):
    def __init__(self, num_channels, stride_level, patch_size_full, dim_tokens, sincos_pos_emb, learnable_pos_emb, image_size):
        super().__init__()
        self.num_channels = num_channels
        self.stride_level = stride_level
        self.patch_size_full = pair(patch_size_full)
        self.dim_tokens = dim_tokens
        self.sincos_pos_emb = sincos_pos_emb
        self.learnable_pos_emb = learnable_pos_emb
        self.image_size = pair(image_size)
        self.num_patches = (self.image_size[0] // patch_size_full) * (self.image_size[1] // patch_size_full)
        self.P_H = max(1, self.patch_size_full[0] // stride_level)
        self.P_W = max(1, self.patch_size_full[1] // stride_level)
        self._pos_cache = None

    def _make_pos_emb(self, dim_tokens):
        h = self.image_size[0] // (self.stride_level * self.P_H)
        w = self.image_size[1] // (self.stride_level * self.P_W)
        if self.sincos_pos_emb:
            self.pos_emb = nn.Parameter(build_2d_sincos_posemb(h=h, w=w, embed_dim=dim_tokens), requires_grad=self.learnable_pos_emb)
        else:
            self.pos_emb = nn.Parameter(torch.zeros(1, dim_tokens, h, w))
            trunc_normal_(self.pos_emb, std=0.02)

    POS_RESIZE_MODE = 'bicubic'

    def pos_table(self, n_h: int, n_w: int) -> torch.Tensor:
        """[n_h*n_w, D] fp32 rows of the positional table (bicubic resize only if the grid differs,
        input_adapters.py:113 -- the identity at the native size; the semantic adapter resizes bilinearly, :321)."""
        key = (n_h, n_w, self.pos_emb.data_ptr(), self.pos_emb._version, self.pos_emb.device)
        if self._pos_cache is None or self._pos_cache[0] != key:
            pe = self.pos_emb.detach()
            if pe.shape[-2:] != (n_h, n_w):
                pe = torch.nn.functional.interpolate(pe, size=(n_h, n_w), mode=self.POS_RESIZE_MODE, align_corners=False)
            self._pos_cache = (key, pe.flatten(2).transpose(1, 2)[0].contiguous())
        return self._pos_cache[1]

    def pos_table_grad(self, n_h: int, n_w: int) -> Optional[torch.Tensor]:
        """the same table ON THE AUTOGRAD TAPE when the positional embedding is learnable (learnable_pos_emb=True, or
        sincos_pos_emb=False whose trunc-normal table is a trained parameter, input_adapters.py:76-87); None otherwise.  The
        fused embedding adds the detached table in its GEMM epilogue; functions.PosEmbGradFn routes the token gradients back
        into this tensor (and through the resize to `pos_emb`)."""
        if not self.pos_emb.requires_grad:
            return None
        pe = self.pos_emb
        if pe.shape[-2:] != (n_h, n_w):
            pe = torch.nn.functional.interpolate(pe, size=(n_h, n_w), mode=self.POS_RESIZE_MODE, align_corners=False)
        return pe.flatten(2).transpose(1, 2)[0]

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb'}


class PatchedInputAdapter(_SpatialAdapterBase):
    """Patchify + project + positional embedding for an image-like modality (input_adapters.py:27-119)."""

    def __init__(self, num_channels: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True, learnable_pos_emb: bool = False,
                 image_size: Union[int, Tuple[int]] = 224):
        super().__init__(num_channels, stride_level, patch_size_full, dim_tokens, sincos_pos_emb, learnable_pos_emb, image_size)
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768):
        self.dim_tokens = dim_tokens
        self._make_pos_emb(dim_tokens)
        self.proj = nn.Conv2d(in_channels=self.num_channels, out_channels=self.dim_tokens,
                              kernel_size=(self.P_H, self.P_W), stride=(self.P_H, self.P_W))

    def grid(self, H, W):
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        assert (H % self.P_H == 0) and (W % self.P_W == 0), \
            f'Image sizes {H}x{W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
        assert self.P_H == self.P_W, 'square patches only'
        return H // self.P_H, W // self.P_W

    def forward(self, x):
        """[B, C, H, W] -> [B, n_h*n_w, D] fp32: every patch embedded (reference semantics)."""
        B, C, H, W = x.shape
        n_h, n_w = self.grid(H, W)
        n = n_h * n_w
        idx = torch.arange(n, dtype=torch.int32, device=x.device)
        meta = dict(B=B, D=self.dim_tokens, P=self.P_H, F=0, nenc=n, idx=[idx], pos=[self.pos_table(n_h, n_w)],
                    pos_fusion=torch.zeros(0, self.dim_tokens, device=x.device))
        X = Fn.EmbedFn.apply(meta, torch.zeros(1, 0, self.dim_tokens, device=x.device), x.float(), self.proj.weight, self.proj.bias)
        return X.view(B, n, self.dim_tokens)


class SemSegInputAdapter(_SpatialAdapterBase):
    """Adapter for semantic class maps (input_adapters.py:209-328; the `dnw` modality of pretrain_mmae_my.py:68-75): a
    learned class embedding per pixel, then Conv2d(dim_class_emb -> D, k = s = P), plus the positional table.  The two are
    folded into one GEMM over one-hot patch rows (functions.EmbedFn, kind "semseg"); parameters, names and shapes are the
    reference's (`class_emb.weight`, `proj.weight`, `proj.bias`, `pos_emb`)."""
    POS_RESIZE_MODE = 'bilinear'
    KIND = 'semseg'

    def __init__(self, num_classes: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens: Optional[int] = None, sincos_pos_emb: int = True, learnable_pos_emb: int = False,
                 image_size: Union[int, Tuple[int]] = 224, dim_class_emb: int = 64, interpolate_class_emb: bool = False,
                 emb_padding_idx: int = None):
        super().__init__(None, stride_level, patch_size_full, dim_tokens, sincos_pos_emb, learnable_pos_emb, image_size)
        del self.num_channels
        self.num_classes = num_classes
        self.dim_class_emb = dim_class_emb
        self.interpolate_class_emb = interpolate_class_emb
        self.emb_padding_idx = emb_padding_idx
        if self.emb_padding_idx is not None:
            self.num_classes += 1
        if interpolate_class_emb:
            raise NotImplementedError("interpolate_class_emb=True is not built (False in the reference's script, pretrain_mmae_my.py:71)")
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768):
        self.dim_tokens = dim_tokens
        self._make_pos_emb(dim_tokens)
        self.class_emb = nn.Embedding(num_embeddings=self.num_classes, embedding_dim=self.dim_class_emb, padding_idx=self.emb_padding_idx)
        trunc_normal_(self.class_emb.weight, std=0.02)
        self.proj = nn.Conv2d(in_channels=self.dim_class_emb, out_channels=self.dim_tokens,
                              kernel_size=(self.P_H, self.P_W), stride=(self.P_H, self.P_W))

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_emb', 'class_emb'}

    def grid(self, H, W):
        assert self.dim_tokens is not None, 'Need to call init(dim_tokens) function first'
        assert (H % self.P_H == 0) and (W % self.P_W == 0), \
            f'Image sizes {H}x{W} must be divisible by patch sizes {self.P_H}x{self.P_W}'
        assert self.P_H == self.P_W, 'square patches only'
        return H // self.P_H, W // self.P_W

    def embed_args(self, x):
        """this modality's tensors for functions.EmbedFn"""
        return [x.to(torch.int64), self.class_emb.weight, self.proj.weight, self.proj.bias]

    def forward(self, x):
        """[B, H, W] int64 class ids -> [B, n_h*n_w, D] fp32: every patch embedded (reference semantics)"""
        B, H, W = x.shape
        if not x.is_cuda:
            raise RuntimeError("SemSegInputAdapter runs on CUDA only (no CPU fallback)")
        n_h, n_w = self.grid(H, W)
        n = n_h * n_w
        idx = torch.arange(n, dtype=torch.int32, device=x.device)
        meta = dict(B=B, D=self.dim_tokens, P=self.P_H, F=0, nenc=n, idx=[idx], pos=[self.pos_table(n_h, n_w)], kinds=["semseg"],
                    padding_idx={0: self.emb_padding_idx}, pos_fusion=torch.zeros(0, self.dim_tokens, device=x.device))
        X = Fn.EmbedFn.apply(meta, torch.zeros(1, 0, self.dim_tokens, device=x.device), *self.embed_args(x))
        return X.view(B, n, self.dim_tokens)


class FusionInputAdapter(_SpatialAdapterBase):
    """Adds the positional table to the learned fusion tokens (input_adapters.py:121-206)."""

    def __init__(self, num_channels: int, stride_level: int, patch_size_full: Union[int, Tuple[int, int]],
                 dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True, learnable_pos_emb: bool = False,
                 image_size: Union[int, Tuple[int]] = 224):
        super().__init__(num_channels, stride_level, patch_size_full, dim_tokens, sincos_pos_emb, learnable_pos_emb, image_size)
        if self.dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768):
        self.dim_tokens = dim_tokens
        self._make_pos_emb(dim_tokens)

    def forward(self, x):
        B, N, C = x.shape
        assert N == self.num_patches
        H, W = self.image_size
        return x + self.pos_table(H // self.P_H, W // self.P_W)[None].to(x.dtype)
