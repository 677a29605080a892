"""Functional restatement of the reference hot path (see ``oracle/__init__.py`` for status).

Everything operates on a flat ``state_dict`` (the reference's key names, Appendix B of SURVEY.md)
plus an :class:`OracleConfig`; there are no ``nn.Module`` objects.  Tensors may require grad, so
``torch.autograd`` gives the reference backward for gradient parity and for the CPU baseline.

``autocast=True`` emulates what CUDA autocast(bf16) does to the reference: Linear / Conv / matmul
operands are cast to bf16 and produce bf16, LayerNorm / softmax / losses run in fp32
(SURVEY.md Appendix A #5, #17).
"""
from __future__ import annotations

import itertools
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

__all__ = [
    "OracleConfig", "S1", "S2", "DEM", "FUSION",
    "sincos_posemb_2d", "init_state_dict", "perturb_state_dict",
    "zorro_layer_norm", "zorro_attention", "geglu_feed_forward", "biased_mlp",
    "zorro_block", "fusion_block", "patch_embed", "semseg_embed", "add_fusion_posemb",
    "sample_alphas", "generate_random_masks", "masks_from_task_masks", "build_input_info",
    "zorro_mask_from_counts", "pool_mask_from_counts",
    "vit_block", "simple_output_adapter", "xattn_output_adapter",
    "multimae_forward", "masked_mse_loss", "masked_l1_loss", "masked_ce_loss", "hard_negative_loss", "dino_loss",
    "pretrain_loss", "count_params",
]

# token type ids: reference zorro_utils.py:14-18
S1, S2, DEM, FUSION = 0, 1, 2, 3


@dataclass
class OracleConfig:
    """Shapes of one model instance.  Mirrors the ctor arguments of the reference's
    ``MultiMAE`` (multimae.py:58-70), ``PatchedInputAdapter`` (input_adapters.py:41-48) and
    ``SpatialOutputAdapter`` (output_adapters_simple.py:60-80)."""
    variant: str = "crossattn"            # 'plain' = multimae.py, 'crossattn' = multimae_crossattn.py
    dim: int = 768
    depth: int = 12
    heads: int = 8
    dim_head: int = 64
    ff_mult: int = 4
    patch: int = 16
    image_size: int = 224
    channels: "OrderedDict[str, int]" = field(
        default_factory=lambda: OrderedDict([("s1", 1), ("s2", 3), ("dem", 1)]))
    return_token_types: Tuple[int, ...] = (S1, S2, DEM, FUSION)
    out_tasks: Tuple[str, ...] = ("s1", "s2", "dem")
    decoder: str = "simple"               # 'simple' = output_adapters_simple.py, 'xattn' = output_adapters.py
    dec_dim: int = 256
    dec_depth: int = 2
    dec_heads: int = 8
    autocast: bool = False                # emulate CUDA autocast(bf16) cast rules

    @property
    def grid(self) -> int:
        return self.image_size // self.patch

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def ff_inner(self) -> int:
        return int(self.dim * self.ff_mult * 2 / 3)   # zorro_utils.py:122


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------

def _lp(t: Optional[torch.Tensor], cfg: OracleConfig):
    """autocast cast of a matmul operand"""
    if t is None or not cfg.autocast:
        return t
    return t.to(torch.bfloat16)


def _linear(x, w, b, cfg: OracleConfig):
    return F.linear(_lp(x, cfg), _lp(w, cfg), _lp(b, cfg))


def _f32(t, cfg: OracleConfig):
    """ops that autocast forces to fp32 (layer_norm, softmax, losses)"""
    return t.float() if cfg.autocast else t


def sincos_posemb_2d(h: int, w: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    """multimae_utils.py:29-45 -> [1, dim, h, w].  Note the (w, h) 'ij' meshgrid then the
    'b (h w) d -> b d h w' reshape (Appendix A #9)."""
    gw = torch.arange(w, dtype=torch.float32)
    gh = torch.arange(h, dtype=torch.float32)
    gw, gh = torch.meshgrid(gw, gh, indexing="ij")
    assert dim % 4 == 0
    pd = dim // 4
    omega = 1.0 / (temperature ** (torch.arange(pd, dtype=torch.float32) / pd))
    ow = gw.flatten()[:, None] * omega[None, :]
    oh = gh.flatten()[:, None] * omega[None, :]
    pe = torch.cat([ow.sin(), ow.cos(), oh.sin(), oh.cos()], dim=1)       # [(h w), dim]
    return pe.reshape(1, h, w, dim).permute(0, 3, 1, 2).contiguous()


def _trunc_normal(shape, std, gen):
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0, generator=gen)
    return t


def _xavier(out_f, in_f, gen, fan_out=None):
    bound = math.sqrt(6.0 / float((fan_out or out_f) + in_f))
    return (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound


def init_state_dict(cfg: OracleConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init weights with the reference's key names/shapes and init *distributions*
    (multimae.py:96-133, multimae_crossattn.py:105-130, output_adapters_simple.py:97-128).  It does
    not replay the reference's RNG stream; golden tests use state_dicts saved from the reference."""
    g = torch.Generator().manual_seed(seed)
    D, gsz, P = cfg.dim, cfg.grid, cfg.patch
    Fn = cfg.num_patches
    inner = cfg.heads * cfg.dim_head
    I = cfg.ff_inner
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def attn(pfx):
        sd[pfx + "norm.gamma"] = torch.ones(D)
        sd[pfx + "norm.beta"] = torch.zeros(D)
        sd[pfx + "to_q.weight"] = _xavier(inner, D, g)
        sd[pfx + "to_kv.weight"] = _xavier(2 * inner, D, g, fan_out=inner)
        sd[pfx + "to_out.weight"] = _xavier(D, inner, g)

    def zblock(pfx):
        sd[pfx + "norm1.gamma"] = torch.ones(D); sd[pfx + "norm1.beta"] = torch.zeros(D)
        attn(pfx + "attn.")
        sd[pfx + "norm2.gamma"] = torch.ones(D); sd[pfx + "norm2.beta"] = torch.zeros(D)
        sd[pfx + "mlp.0.gamma"] = torch.ones(D); sd[pfx + "mlp.0.beta"] = torch.zeros(D)
        sd[pfx + "mlp.1.weight"] = _xavier(2 * I, D, g)
        sd[pfx + "mlp.3.weight"] = _xavier(D, I, g)

    sd["return_tokens"] = _trunc_normal((1, len(cfg.return_token_types), D), 0.02, g)
    sd["fusion_tokens"] = _trunc_normal((1, Fn, D), 0.02, g)
    if cfg.variant == "crossattn":
        for t in ("s1", "s2", "dem"):
            sd[f"return_token_{t}"] = torch.randn(1, 1, D, generator=g)
        sd["mask_embedding"] = torch.zeros(1, Fn, D)
    for t, C in cfg.channels.items():
        sd[f"input_adapters.{t}.pos_emb"] = sincos_posemb_2d(gsz, gsz, D)
        sd[f"input_adapters.{t}.proj.weight"] = _xavier(D, C * P * P, g).reshape(D, C, P, P)
        sd[f"input_adapters.{t}.proj.bias"] = (torch.rand(D, generator=g) * 2 - 1) / math.sqrt(C * P * P)
    sd["input_adapters.fusion.pos_emb"] = sincos_posemb_2d(gsz, gsz, D)
    attn("attn_pool.")
    sd["mlp.fc1.weight"] = _xavier(4 * D, D, g); sd["mlp.fc1.bias"] = torch.zeros(4 * D)
    sd["mlp.fc2.weight"] = _xavier(D, 4 * D, g); sd["mlp.fc2.bias"] = torch.zeros(D)
    for i in range(cfg.depth):
        zblock(f"blocks.{i}.")
        if cfg.variant == "crossattn":
            zblock(f"fus_blocks.{i}.")
    sd["norm.gamma"] = torch.ones(D); sd["norm.beta"] = torch.zeros(D)

    d = cfg.dec_dim
    for t in cfg.out_tasks:
        p = f"output_adapters.{t}."
        C = cfg.channels[t]
        sd[p + "pos_emb"] = sincos_posemb_2d(gsz, gsz, d)
        for c in cfg.channels:
            sd[p + f"task_embeddings.{c}"] = _trunc_normal((1, 1, d), 0.02, g)
        if cfg.decoder == "xattn":
            sd[p + "mask_token"] = torch.zeros(1, 1, d)
            sd[p + "decoder.q.weight"] = _xavier(d, d, g); sd[p + "decoder.q.bias"] = torch.zeros(d)
            sd[p + "decoder.kv.weight"] = _xavier(2 * d, d, g, fan_out=d); sd[p + "decoder.kv.bias"] = torch.zeros(2 * d)
            sd[p + "decoder.proj.weight"] = _xavier(d, d, g); sd[p + "decoder.proj.bias"] = torch.zeros(d)
            for n in ("context_norm", "query_norm", "out_norm"):
                sd[p + n + ".weight"] = torch.ones(d); sd[p + n + ".bias"] = torch.zeros(d)
            sd[p + "mlp.fc1.weight"] = _xavier(4 * d, d, g); sd[p + "mlp.fc1.bias"] = torch.zeros(4 * d)
            sd[p + "mlp.fc2.weight"] = _xavier(d, 4 * d, g); sd[p + "mlp.fc2.bias"] = torch.zeros(d)
        for j in range(cfg.dec_depth):
            q = p + f"decoder_transformer.{j}."
            sd[q + "norm1.weight"] = torch.ones(d); sd[q + "norm1.bias"] = torch.zeros(d)
            sd[q + "attn.qkv.weight"] = _xavier(3 * d, d, g, fan_out=d); sd[q + "attn.qkv.bias"] = torch.zeros(3 * d)
            sd[q + "attn.proj.weight"] = _xavier(d, d, g); sd[q + "attn.proj.bias"] = torch.zeros(d)
            sd[q + "norm2.weight"] = torch.ones(d); sd[q + "norm2.bias"] = torch.zeros(d)
            sd[q + "mlp.fc1.weight"] = _xavier(4 * d, d, g); sd[q + "mlp.fc1.bias"] = torch.zeros(4 * d)
            sd[q + "mlp.fc2.weight"] = _xavier(d, 4 * d, g); sd[q + "mlp.fc2.bias"] = torch.zeros(d)
        sd[p + "out_proj.weight"] = _xavier(C * P * P, d, g); sd[p + "out_proj.bias"] = torch.zeros(C * P * P)
        sd[p + "proj_context.weight"] = _xavier(d, D, g); sd[p + "proj_context.bias"] = torch.zeros(d)
    return sd


def perturb_state_dict(sd, seed: int = 7, scale: float = 0.05):
    """Deterministically move every trainable tensor off its init value (gammas off 1, biases and
    ``mask_embedding`` off 0) so parity tests exercise every term.  Buffers/pos-embs are untouched."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith(".beta") or k.endswith("pos_emb"):
            out[k] = v.clone()
        else:
            out[k] = v + scale * torch.randn(v.shape, generator=g) * (1.0 if v.dim() <= 1 or "token" in k or "embedding" in k else 0.2)
    return out


def count_params(sd) -> int:
    return sum(v.numel() for k, v in sd.items() if not k.endswith(".beta"))


# ----------------------------------------------------------------------------------------------
# encoder primitives (zorro_utils.py)
# ----------------------------------------------------------------------------------------------

def zorro_layer_norm(x, gamma, cfg: OracleConfig, eps: float = 1e-5):
    """zorro_utils.py:103-110: F.layer_norm with learnable gamma and a zero beta buffer."""
    x = _f32(x, cfg)
    return F.layer_norm(x, x.shape[-1:], gamma, torch.zeros_like(gamma), eps)


def zorro_attention(sd, pfx: str, x, cfg: OracleConfig, context=None, attn_mask=None):
    """zorro_utils.py:170-194.  LN on the query side only; q pre-scaled; masked_fill(-finfo.max)
    (so an all-masked row softmaxes to uniform, Appendix A #4); no biases."""
    h, dh = cfg.heads, cfg.dim_head
    x = zorro_layer_norm(x, sd[pfx + "norm.gamma"], cfg)
    kv_x = x if context is None else context
    q = _linear(x, sd[pfx + "to_q.weight"], None, cfg)
    kv = _linear(kv_x, sd[pfx + "to_kv.weight"], None, cfg)
    k, v = kv.chunk(2, dim=-1)
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    q = q.reshape(B, Nq, h, dh).transpose(1, 2)
    k = k.reshape(B, Nk, h, dh).transpose(1, 2)
    v = v.reshape(B, Nk, h, dh).transpose(1, 2)
    q = q * (dh ** -0.5)
    sim = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        sim = sim.masked_fill(~attn_mask, -torch.finfo(sim.dtype).max)
    attn = _f32(sim, cfg).softmax(dim=-1)
    out = _lp(attn, cfg) @ v if cfg.autocast else attn @ v
    out = out.transpose(1, 2).reshape(B, Nq, h * dh)
    return _linear(out, sd[pfx + "to_out.weight"], None, cfg)


def geglu_feed_forward(sd, pfx: str, x, cfg: OracleConfig):
    """zorro_utils.py:115-128: LN -> Linear(D, 2I) -> gelu(gate) * value -> Linear(I, D).
    First half of the 2I columns is the value, second half the gate; exact-erf GELU."""
    x = zorro_layer_norm(x, sd[pfx + "0.gamma"], cfg)
    u = _linear(x, sd[pfx + "1.weight"], None, cfg)
    val, gate = u.chunk(2, dim=-1)
    return _linear(F.gelu(gate) * val, sd[pfx + "3.weight"], None, cfg)


def biased_mlp(sd, pfx: str, x, cfg: OracleConfig):
    """zorro_utils.py:131-148 / multimae_utils.py:138-155: fc1 -> GELU -> fc2 (with biases)."""
    x = _linear(x, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"], cfg)
    x = F.gelu(x)
    return _linear(x, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"], cfg)


def zorro_block(sd, pfx: str, x, attn_mask, cfg: OracleConfig):
    """zorro_utils.py:237-240 (DropPath rate 0 -> identity).  Note the double LayerNorm."""
    x = x + zorro_attention(sd, pfx + "attn.", zorro_layer_norm(x, sd[pfx + "norm1.gamma"], cfg), cfg,
                            attn_mask=attn_mask)
    x = x + geglu_feed_forward(sd, pfx + "mlp.", zorro_layer_norm(x, sd[pfx + "norm2.gamma"], cfg), cfg)
    return x


def fusion_block(sd, pfx: str, slots, cfg: OracleConfig):
    """Working ``Block_Fusion``: downstream/instance_segmentation/modeling/multimae/zorro_utils.py:243-258.
    slots [B, F, S, D]: unmasked self-attention over the S slots of every position, keep the last
    (fusion) slot, then the GEGLU FFN."""
    B, Fn, S, D = slots.shape
    x = slots.reshape(B * Fn, S, D)
    x = x + zorro_attention(sd, pfx + "attn.", zorro_layer_norm(x, sd[pfx + "norm1.gamma"], cfg), cfg)
    x = x[:, -1, :].reshape(B, Fn, D)
    x = x + geglu_feed_forward(sd, pfx + "mlp.", zorro_layer_norm(x, sd[pfx + "norm2.gamma"], cfg), cfg)
    return x


# ----------------------------------------------------------------------------------------------
# adapters (input_adapters.py)
# ----------------------------------------------------------------------------------------------

def patch_embed(sd, pfx: str, img, cfg: OracleConfig):
    """PatchedInputAdapter.forward, input_adapters.py:97-119: Conv2d(k=s=P)+bias, flatten to
    [B, (nh nw), D], add the sin-cos pos-emb (bicubic resize is the identity at native size)."""
    P = cfg.patch
    B, C, H, W = img.shape
    assert H % P == 0 and W % P == 0
    y = F.conv2d(_lp(img, cfg), _lp(sd[pfx + "proj.weight"], cfg), _lp(sd[pfx + "proj.bias"], cfg), stride=P)
    y = y.flatten(2).transpose(1, 2)
    pe = sd[pfx + "pos_emb"]
    if pe.shape[-2:] != (H // P, W // P):
        pe = F.interpolate(pe, size=(H // P, W // P), mode="bicubic", align_corners=False)
    return y + pe.flatten(2).transpose(1, 2)


def semseg_embed(sd, pfx: str, x, patch: int, padding_idx=None, lowp=None):
    """SemSegInputAdapter.forward, input_adapters.py:299-328 (interpolate_class_emb=False): class embedding per pixel,
    Conv2d(k = s = P) + bias, flatten, + the positional table (bilinear resize if the grid differs).  x: [B, H, W] int64."""
    emb = F.embedding(x, sd[pfx + "class_emb.weight"], padding_idx=padding_idx)            # [B, H, W, E]
    y = F.conv2d(emb.permute(0, 3, 1, 2), sd[pfx + "proj.weight"], sd[pfx + "proj.bias"], stride=patch)
    y = y.flatten(2).transpose(1, 2)
    pe = sd[pfx + "pos_emb"]
    nh, nw = x.shape[1] // patch, x.shape[2] // patch
    if pe.shape[-2:] != (nh, nw):
        pe = F.interpolate(pe, size=(nh, nw), mode="bilinear")
    return y + pe.flatten(2).transpose(1, 2)


def add_fusion_posemb(sd, fusion_tokens):
    """FusionInputAdapter.forward, input_adapters.py:185-206."""
    pe = sd["input_adapters.fusion.pos_emb"]
    assert fusion_tokens.shape[1] == pe.shape[-1] * pe.shape[-2]
    return fusion_tokens + pe.flatten(2).transpose(1, 2)


# ----------------------------------------------------------------------------------------------
# masking (multimae.py:165-255, 287-306, 372-383, 410-426)
# ----------------------------------------------------------------------------------------------

def sample_alphas(n_tasks: int, alphas: Sequence[float], eps: float = 1e-5):
    """multimae.py:165-180 with B=1 (the reference's 'same mask for the batch' edit, :204)."""
    choices = torch.Tensor([list(i) for i in itertools.product([0, 1], repeat=n_tasks)][1:])
    pick = torch.randint(0, len(choices), (1,))
    return torch.index_select(choices, 0, pick) * torch.tensor(alphas) + eps


def generate_random_masks(num_tokens_per_task: "OrderedDict[str, int]", batch: int, num_encoded_tokens: int,
                          device, alphas=1.0, sample_tasks_uniformly: bool = False):
    """multimae.py:182-255.  Consumes the global torch RNG in the reference's order: (randint +)
    Dirichlet on the CPU, one ``torch.rand(1, n)`` per task on ``device``, one ``rand_like``."""
    from torch.distributions.dirichlet import Dirichlet
    n_tasks = len(num_tokens_per_task)
    alphas = [alphas] * n_tasks if isinstance(alphas, float) else list(alphas)
    if sample_tasks_uniformly:
        dist = Dirichlet(sample_alphas(n_tasks, alphas)).sample().to(device)
    else:
        dist = Dirichlet(torch.Tensor(alphas)).sample((1,)).to(device)
    samples_per_task = (dist * num_encoded_tokens).round().long()
    masks = []
    for i, n in enumerate(num_tokens_per_task.values()):
        noise = torch.rand(1, n, device=device)
        order = torch.argsort(noise, dim=1)
        ranks = torch.gather(torch.arange(n, device=device).unsqueeze(0), 1, order)
        masks.append(torch.where(ranks < samples_per_task[:, i].unsqueeze(1), 0, 1))
    mask_all = torch.cat(masks, dim=1)
    ids_shuffle = torch.argsort(mask_all + torch.rand_like(mask_all.float()), dim=1)
    ids_restore = torch.argsort(ids_shuffle, dim=1)
    ids_keep = ids_shuffle[:, :num_encoded_tokens]
    mask_all = torch.ones_like(mask_all)
    mask_all[:, :num_encoded_tokens] = 0
    mask_all = torch.gather(mask_all, 1, ids_restore)
    parts = torch.split(mask_all, list(num_tokens_per_task.values()), dim=1)
    task_masks = {t: m.repeat(batch, 1) for t, m in zip(num_tokens_per_task, parts)}
    return task_masks, ids_keep.repeat(batch, 1), ids_restore.repeat(batch, 1)


def masks_from_task_masks(task_masks: Dict[str, torch.Tensor], task_order: Sequence[str]):
    """explicit-mask branch, multimae.py:372-376 (the ``.sum()`` is over the whole tensor)."""
    mask_all = torch.cat([task_masks[t] for t in task_order], dim=1)
    ids_shuffle = torch.argsort(mask_all, dim=1, stable=True)
    ids_restore = torch.argsort(ids_shuffle, dim=1, stable=True)
    ids_keep = ids_shuffle[:, :int((mask_all == 0).sum())]
    return ids_keep, ids_restore


def build_input_info(num_tokens_per_task: "OrderedDict[str, int]", image_size):
    """multimae.py:287-306."""
    info = OrderedDict()
    info["tasks"] = {}
    i = 0
    for t, n in num_tokens_per_task.items():
        info["tasks"][t] = {"num_tokens": n, "has_2d_posemb": True, "start_idx": i, "end_idx": i + n}
        i += n
    info["image_size"] = image_size
    info["num_task_tokens"] = i
    return info


def _token_types(counts: Sequence[int], n_fusion: int, device):
    types = []
    for t, n in zip((S1, S2, DEM), counts):
        types += [t] * int(n)
    types += [FUSION] * n_fusion
    return torch.tensor(types, dtype=torch.long, device=device)


def zorro_mask_from_counts(counts, n_fusion, device):
    """multimae.py:410-426: same type attends to same type, fusion attends to everything."""
    tt = _token_types(counts, n_fusion, device)
    return (tt[:, None] == tt[None, :]) | (tt[:, None] == FUSION)


def pool_mask_from_counts(return_types: Sequence[int], counts, n_fusion, device):
    """multimae.py:447-451."""
    tt = _token_types(counts, n_fusion, device)
    rt = torch.tensor(list(return_types), dtype=torch.long, device=device)
    return (rt[:, None] == tt[None, :]) | (rt[:, None] == FUSION)


# ----------------------------------------------------------------------------------------------
# decoders (output_adapters_simple.py, output_adapters.py, multimae_utils.py)
# ----------------------------------------------------------------------------------------------

def _nn_layer_norm(sd, pfx, x, cfg, eps=1e-6):
    x = _f32(x, cfg)
    return F.layer_norm(x, x.shape[-1:], sd[pfx + "weight"], sd[pfx + "bias"], eps)


def _vit_attention(sd, pfx, x, heads, cfg):
    """multimae_utils.py:158-182 (scale applied after QK^T)."""
    B, N, C = x.shape
    qkv = _linear(x, sd[pfx + "qkv.weight"], sd[pfx + "qkv.bias"], cfg)
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    attn = _f32(attn, cfg).softmax(dim=-1)
    x = ((_lp(attn, cfg) if cfg.autocast else attn) @ v).transpose(1, 2).reshape(B, N, C)
    return _linear(x, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"], cfg)


def _vit_cross_attention(sd, pfx, x, context, heads, cfg):
    """multimae_utils.py:185-214."""
    B, N, C = x.shape
    M = context.shape[1]
    q = _linear(x, sd[pfx + "q.weight"], sd[pfx + "q.bias"], cfg).reshape(B, N, heads, C // heads).permute(0, 2, 1, 3)
    kv = _linear(context, sd[pfx + "kv.weight"], sd[pfx + "kv.bias"], cfg)
    kv = kv.reshape(B, M, 2, heads, C // heads).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    attn = (q @ k.transpose(-2, -1)) * ((C // heads) ** -0.5)
    attn = _f32(attn, cfg).softmax(dim=-1)
    x = ((_lp(attn, cfg) if cfg.autocast else attn) @ v).transpose(1, 2).reshape(B, N, -1)
    return _linear(x, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"], cfg)


def vit_block(sd, pfx: str, x, heads: int, cfg: OracleConfig):
    """multimae_utils.py:217-232: pre-LN ViT block, nn.LayerNorm(eps=1e-6), biased linears."""
    x = x + _vit_attention(sd, pfx + "attn.", _nn_layer_norm(sd, pfx + "norm1.", x, cfg), heads, cfg)
    x = x + biased_mlp(sd, pfx + "mlp.", _nn_layer_norm(sd, pfx + "norm2.", x, cfg), cfg)
    return x


def _unpatchify(x, cfg: OracleConfig, C: int, H: int, W: int):
    """'b (nh nw) (c ph pw) -> b c (nh ph) (nw pw)', output_adapters_simple.py:183-186."""
    P = cfg.patch
    B = x.shape[0]
    nh, nw = H // P, W // P
    return x.reshape(B, nh, nw, C, P, P).permute(0, 3, 1, 4, 2, 5).reshape(B, C, H, W)


def simple_output_adapter(sd, task: str, enc_tokens, image_size, cfg: OracleConfig):
    """output_adapters_simple.py:146-188."""
    p = f"output_adapters.{task}."
    H, W = image_size
    x = _linear(enc_tokens, sd[p + "proj_context.weight"], sd[p + "proj_context.bias"], cfg)
    key = p + f"task_embeddings.{task}"
    if key in sd:
        x = x + sd[key]
    for j in range(cfg.dec_depth):
        x = vit_block(sd, p + f"decoder_transformer.{j}.", x, cfg.dec_heads, cfg)
    x = _linear(x, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"], cfg)
    return _unpatchify(x, cfg, cfg.channels[task], H, W)


def xattn_output_adapter(sd, task: str, enc_tokens, input_info, ids_keep, ids_restore, cfg: OracleConfig):
    """output_adapters.py:160-282 (use_task_queries=True, use_xattn=True)."""
    p = f"output_adapters.{task}."
    H, W = input_info["image_size"]
    B = enc_tokens.shape[0]
    d = cfg.dec_dim
    ctx = _linear(enc_tokens, sd[p + "proj_context.weight"], sd[p + "proj_context.bias"], cfg)
    n_mask = input_info["num_task_tokens"] - ctx.shape[1]
    full = torch.cat([ctx, sd[p + "mask_token"].expand(B, n_mask, d).to(ctx.dtype)], dim=1)
    full = torch.gather(full, 1, ids_restore.unsqueeze(-1).expand(-1, -1, d))
    pe = sd[p + "pos_emb"]
    nh, nw = H // cfg.patch, W // cfg.patch
    if pe.shape[-2:] != (nh, nw):
        pe = F.interpolate(pe, size=(nh, nw), mode="bilinear", align_corners=False)
    pe = pe.flatten(2).transpose(1, 2)
    embs = []
    for t, info in input_info["tasks"].items():
        key = p + f"task_embeddings.{t}"
        e = sd[key].expand(B, info["num_tokens"], d) if key in sd else torch.zeros(B, info["num_tokens"], d, device=ctx.device)
        embs.append(e + pe)
    full = full + torch.cat(embs, dim=1)
    ti = input_info["tasks"][task]
    queries = full[:, ti["start_idx"]:ti["end_idx"]]
    context = torch.gather(full, 1, ids_keep.unsqueeze(-1).expand(-1, -1, d))
    x = _vit_cross_attention(sd, p + "decoder.", _nn_layer_norm(sd, p + "query_norm.", queries, cfg),
                             _nn_layer_norm(sd, p + "context_norm.", context, cfg), cfg.dec_heads, cfg)
    x = x + biased_mlp(sd, p + "mlp.", _nn_layer_norm(sd, p + "out_norm.", x, cfg), cfg)
    for j in range(cfg.dec_depth):
        x = vit_block(sd, p + f"decoder_transformer.{j}.", x, cfg.dec_heads, cfg)
    x = _linear(x, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"], cfg)
    return _unpatchify(x, cfg, cfg.channels[task], H, W)


# ----------------------------------------------------------------------------------------------
# the model forward (multimae.py:308-487, multimae_crossattn.py:331-545)
# ----------------------------------------------------------------------------------------------

def multimae_forward(sd, cfg: OracleConfig, x: Dict[str, torch.Tensor], mask_inputs: bool = True,
                     task_masks: Optional[Dict[str, torch.Tensor]] = None, num_encoded_tokens: int = 128,
                     alphas=1.0, sample_tasks_uniformly: bool = False, decode: bool = True,
                     return_token_indices=None, return_aux: bool = False):
    """Returns the reference's tuple: (preds, task_masks, return_tokens, ori_tokens,
    encoder_fusion_tokens[, return_token_s1, _s2, _dem]) -- or (tokens, return_tokens, task_masks)
    when ``decode`` is False (the ``output_adapters is None`` branch, multimae.py:458-459)."""
    dev = x["s1"].device
    B, _, H, W = x["s1"].shape
    Fn = cfg.num_patches

    tok = OrderedDict((t, patch_embed(sd, f"input_adapters.{t}.", img, cfg)) for t, img in x.items()
                      if t in cfg.channels)
    fusion = add_fusion_posemb(sd, sd["fusion_tokens"].expand(B, -1, -1))
    n_per_task = OrderedDict((t, v.shape[1]) for t, v in tok.items())
    input_info = build_input_info(n_per_task, (H, W))
    if not mask_inputs:
        num_encoded_tokens = sum(n_per_task.values())
    if task_masks is None:
        task_masks, ids_keep, ids_restore = generate_random_masks(
            n_per_task, B, num_encoded_tokens, dev, alphas=alphas, sample_tasks_uniformly=sample_tasks_uniformly)
    else:
        ids_keep, ids_restore = masks_from_task_masks(task_masks, list(tok.keys()))

    idx = {t: (task_masks[t][0] == 0).nonzero(as_tuple=True)[0] for t in ("s1", "s2", "dem")}
    counts = [len(idx[t]) for t in ("s1", "s2", "dem")]
    tokens = torch.cat([tok[t][:, idx[t]] for t in ("s1", "s2", "dem")] + [fusion], dim=1)
    nenc = num_encoded_tokens
    zmask = zorro_mask_from_counts(counts, Fn, dev)

    for i in range(cfg.depth):
        if cfg.variant == "crossattn":
            # multimae_crossattn.py:450-470
            me = sd["mask_embedding"].expand(B, -1, -1)
            offs = [0, counts[0], nenc - counts[2]]
            slots = []
            for t, o, n in zip(("s1", "s2", "dem"), offs, counts):
                s = me.clone().to(tokens.dtype)
                s[:, idx[t]] = tokens[:, o:o + n]
                slots.append(s)
            slots.append(tokens[:, nenc:])
            fus = fusion_block(sd, f"fus_blocks.{i}.", torch.stack(slots, dim=2), cfg)
            tokens = torch.cat([tokens[:, :nenc], fus], dim=1)
        tokens = zorro_block(sd, f"blocks.{i}.", tokens, zmask, cfg)
    tokens = zorro_layer_norm(tokens, sd["norm.gamma"], cfg)

    rt = sd["return_tokens"]
    rtypes = list(cfg.return_token_types)
    if return_token_indices is not None:
        rt = rt[:, list(return_token_indices)]
        rtypes = [rtypes[i] for i in return_token_indices]
    pmask = pool_mask_from_counts(rtypes, counts, Fn, dev)

    def pool(query, context, mask):
        r = zorro_attention(sd, "attn_pool.", query.expand(B, -1, -1), cfg, context=context, attn_mask=mask)
        return r + biased_mlp(sd, "mlp.", zorro_layer_norm(r, sd["norm.gamma"], cfg), cfg)

    return_tokens = pool(rt, tokens, pmask)
    if not decode:
        return tokens, return_tokens, task_masks

    ori_tokens = tokens[:, :nenc]
    enc_fusion = tokens[:, nenc:]
    preds = {}
    for t in cfg.out_tasks:
        if cfg.decoder == "simple":
            preds[t] = simple_output_adapter(sd, t, enc_fusion, (H, W), cfg)
        else:
            preds[t] = xattn_output_adapter(sd, t, enc_fusion, input_info, ids_keep, ids_restore, cfg)
    out = (preds, task_masks, return_tokens, ori_tokens, enc_fusion)
    if cfg.variant == "crossattn":
        # multimae_crossattn.py:529-543
        out = out + tuple(pool(sd[f"return_token_{t}"], enc_fusion[:, idx[t]], None) for t in ("s1", "s2", "dem"))
    if return_aux:
        return out, {"ids_keep": ids_keep, "ids_restore": ids_restore, "idx": idx, "counts": counts,
                     "input_info": input_info}
    return out


# ----------------------------------------------------------------------------------------------
# losses (criterion.py)
# ----------------------------------------------------------------------------------------------

def _masked_recon(per_pixel, mask, patch: int):
    """criterion.py:99-113 (shared tail of MaskedMSELoss / MaskedL1Loss)."""
    B, C, H, W = per_pixel.shape
    if mask is None:
        return per_pixel.mean()
    if mask.sum() == 0:
        return torch.tensor(0).to(per_pixel.device)
    nh, nw = H // patch, W // patch
    m = mask.reshape(B, nh, nw).float()
    m = F.interpolate(m.unsqueeze(1), size=(H, W), mode="nearest").squeeze(1)
    loss = per_pixel.mean(dim=1) * m
    loss = loss.flatten(1).sum(dim=1) / m.flatten(1).sum(dim=1)
    return loss.nanmean()


def masked_mse_loss(pred, target, mask=None, patch: int = 16):
    """MaskedMSELoss.forward (norm_pix=False), criterion.py:85-115."""
    return _masked_recon(F.mse_loss(pred, target, reduction="none"), mask, patch)


def masked_l1_loss(pred, target, mask=None, patch: int = 16):
    """MaskedL1Loss.forward (norm_pix=False), criterion.py:142-172."""
    return _masked_recon(F.l1_loss(pred, target, reduction="none"), mask, patch)


def masked_ce_loss(logits, target, mask=None, patch: int = 16):
    """MaskedCrossEntropyLoss.forward, criterion.py:37-58 (label_smoothing = 0): per-pixel cross-entropy over the class
    axis, patch mask nearest-upsampled to pixels, per-sample sum / mask sum, batch nanmean; an all-zero mask returns 0."""
    loss = F.cross_entropy(logits.float(), target, reduction="none")
    if mask is None:
        return loss.mean()
    if mask.sum() == 0:
        return torch.zeros((), device=loss.device)
    H, W = logits.shape[-2:]
    nh, nw = H // patch, W // patch
    m = mask.reshape(mask.shape[0], nh, nw).float()
    m = F.interpolate(m.unsqueeze(1), size=(H, W), mode="nearest").squeeze(1)
    loss = (loss * m).flatten(1).sum(1) / m.flatten(1).sum(1)
    return loss.nanmean()


def hard_negative_loss(out_1, out_2, tau_plus=0.1, beta=1.0, temperature=0.5):
    """HardNegtive_loss.forward ('hard' estimator), criterion.py:233-268; device-agnostic (the
    reference hard-codes ``.cuda()`` at :242 and builds the mask in a Python loop :224-231)."""
    B = out_1.shape[0]
    out_1 = F.normalize(out_1, dim=1)
    out_2 = F.normalize(out_2, dim=1)
    out = torch.cat([out_1, out_2], dim=0)
    neg = torch.exp(out @ out.t() / temperature)
    eye = torch.eye(B, dtype=torch.bool, device=out.device)
    keep = ~torch.cat([torch.cat([eye, eye], 1), torch.cat([eye, eye], 1)], 0)
    neg = neg.masked_select(keep).view(2 * B, -1)
    pos = torch.exp(torch.sum(out_1 * out_2, dim=-1) / temperature)
    pos = torch.cat([pos, pos], dim=0)
    N = B * 2 - 2
    imp = (beta * neg.log()).exp()
    reweight = (imp * neg).sum(dim=-1) / imp.mean(dim=-1)
    Ng = (-tau_plus * N * pos + reweight) / (1 - tau_plus)
    Ng = torch.clamp(Ng, min=N * math.e ** (-1 / temperature))
    return (-torch.log(pos / (pos + Ng))).mean()


def dino_loss(student, teacher, teacher_temp=0.04, student_temp=0.1):
    """dino_loss_func, criterion.py:328-335."""
    s = F.log_softmax(F.normalize(student, dim=1) / student_temp, dim=-1)
    t = F.softmax(F.normalize(teacher, dim=1) / teacher_temp, dim=-1).detach()
    return torch.sum(-t * s, dim=-1).mean()


def pretrain_loss(out, targets: Dict[str, torch.Tensor], cfg: OracleConfig, contrastive_weight: float = 0.3):
    """Loss assembly of ``train_one_epoch``, pretrain_mmae.py:476-500: masked MSE for s1/s2, masked
    L1 for dem (DOMAIN_CONF :45-72), + 0.3 * sum of three DINO-style terms (crossattn variant)."""
    preds, masks = out[0], out[1]
    losses = {}
    for t, p in preds.items():
        fn = masked_l1_loss if t == "dem" else masked_mse_loss
        losses[t] = fn(p.float(), targets[t], masks.get(t), cfg.patch)
    total = sum(losses.values())
    if len(out) == 8:
        pooled = out[2].float()
        feats = [f.squeeze(1) for f in torch.chunk(pooled, pooled.shape[1], dim=1)]
        toks = [o.float().squeeze(1) for o in out[5:8]]
        contra = sum(dino_loss(toks[i], feats[i]) for i in range(3))
        losses["contrastive"] = contra
        total = total + contrastive_weight * contra
    return total, losses
