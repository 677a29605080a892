"""CPU oracle (TEST INFRASTRUCTURE) for the downstream incomplete-modality backbone `ViTBaseline`:
downstream/instance_segmentation/modeling/multimae/multimae_big_imcomplete.py:418-454 (heads), :534-659
(forward_features) and :661-676 (forward), eval mode (every token of every PRESENT modality is encoded).
Pinned by tests/golden/vitbaseline.pt, produced by the reference's own class (tests/golden/make_golden.py)."""
from collections import OrderedDict
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from .functional import (OracleConfig, add_fusion_posemb, fusion_block, init_state_dict, patch_embed, zorro_block,
                         zorro_layer_norm)

FUSION = 3
TYPE = {"s1": 0, "s2": 1, "dem": 2}


def vit_baseline_state_dict(cfg: OracleConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of the downstream model: the fusion-block encoder's keys (without the per-modality return tokens of the
    pre-training variant) plus the pyramid heads up1 (ConvT, GroupNorm, GELU, ConvT) and up2 (ConvT)."""
    sd = OrderedDict((k, v) for k, v in init_state_dict(cfg, seed=seed).items()
                     if not k.startswith("return_token_") and not k.startswith("output_adapters."))
    g = torch.Generator().manual_seed(seed + 101)
    D = cfg.dim
    r = lambda *s: torch.randn(*s, generator=g) * 0.05
    sd["up1.0.weight"], sd["up1.0.bias"] = r(D, D, 2, 2), r(D)
    sd["up1.1.weight"], sd["up1.1.bias"] = 1.0 + r(D), r(D)
    sd["up1.3.weight"], sd["up1.3.bias"] = r(D, D, 2, 2), r(D)
    sd["up2.weight"], sd["up2.bias"] = r(D, D, 2, 2), r(D)
    return sd


def vit_baseline_flags(depth: int):
    return [i for i in range(-1, depth, depth // 4)][1:]          # (:431)


def vit_baseline_forward(sd, cfg: OracleConfig, x: Dict[str, torch.Tensor], in_domains: Sequence[str] = ("s1", "s2", "dem"),
                         return_taps: bool = False):
    """eval-mode forward: [f1, f2, f3, f4] feature maps.  Absent modalities contribute neither tokens nor a slot in the
    per-position modality attention (:586-606, :633-647)."""
    present = [t for t in x if t in in_domains]
    B, _, H, W = x[present[0]].shape
    Fn = cfg.num_patches
    tok = OrderedDict((t, patch_embed(sd, f"input_adapters.{t}.", x[t], cfg)) for t in present)
    fusion = add_fusion_posemb(sd, sd["fusion_tokens"].expand(B, -1, -1))
    tokens = torch.cat([tok[t] for t in present] + [fusion], dim=1)        # every token is visible in eval mode
    nenc = Fn * len(present)
    types = torch.tensor(sum(([TYPE[t]] * Fn for t in present), []) + [FUSION] * Fn)
    zmask = (types[:, None] == types[None, :]) | (types[:, None] == FUSION)          # (:609-628)
    flags = vit_baseline_flags(cfg.depth)
    outs = []
    for i in range(cfg.depth):
        slots = [tokens[:, m * Fn:(m + 1) * Fn] for m in range(len(present))] + [tokens[:, nenc:]]
        fus = fusion_block(sd, f"fus_blocks.{i}.", torch.stack(slots, dim=2), cfg)
        tokens = torch.cat([tokens[:, :nenc], fus], dim=1)
        tokens = zorro_block(sd, f"blocks.{i}.", tokens, zmask, cfg)
        if i in flags:
            outs.append(tokens[:, nenc:])
    if return_taps:          # forward_features' own return value: the un-normalised fusion tokens after the flagged blocks
        return outs
    nh, nw = H // cfg.patch, W // cfg.patch
    maps = [zorro_layer_norm(f, sd["norm.gamma"], cfg).transpose(1, 2).reshape(B, cfg.dim, nh, nw) for f in outs]
    f1 = F.conv_transpose2d(maps[0], sd["up1.0.weight"], sd["up1.0.bias"], stride=2)
    f1 = F.gelu(F.group_norm(f1, 32, sd["up1.1.weight"], sd["up1.1.bias"]))
    f1 = F.conv_transpose2d(f1, sd["up1.3.weight"], sd["up1.3.bias"], stride=2)
    f2 = F.conv_transpose2d(maps[1], sd["up2.weight"], sd["up2.bias"], stride=2)
    return [f1, f2, maps[2], F.max_pool2d(maps[3], 2, 2)]
