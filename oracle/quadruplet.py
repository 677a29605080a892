"""CPU oracle (TEST INFRASTRUCTURE) for the 4-modality variant: pretraining/multimae/multimae_quadruplet.py:308-491 with the
semantic `dnw` input (SemSegInputAdapter, input_adapters.py:209-328) and the loss assembly of pretrain_mmae_my.py
(MSE s1 / s2, L1 dem, masked cross-entropy dnw).  Pinned by tests/golden/quadruplet.pt (the reference's own model)."""
from collections import OrderedDict
from dataclasses import replace
from typing import Dict

import torch

from .functional import (OracleConfig, _trunc_normal, _xavier, add_fusion_posemb, biased_mlp, build_input_info,
                         generate_random_masks, init_state_dict, masked_ce_loss, masked_l1_loss, masked_mse_loss,
                         masks_from_task_masks, patch_embed, semseg_embed, simple_output_adapter, sincos_posemb_2d,
                         zorro_attention, zorro_block, zorro_layer_norm)

TASKS = ("s1", "s2", "dem", "dnw")
TYPE = {"s1": 0, "s2": 1, "dem": 2, "dnw": 3}
FUSION = 4          # zorro_utils_quadruplet.py:18-23
NUM_CLASSES = 9     # pretrain_mmae_my.py:68-75


def quad_config(**kw) -> OracleConfig:
    """channels follow pretrain_mmae_my.py:45-75; for `dnw` the entry is the number of classes = channels of its decoder"""
    cfg = OracleConfig(variant="plain", decoder="simple", **kw)
    return replace(cfg, channels=OrderedDict([("s1", 2), ("s2", 4), ("dem", 1), ("dnw", NUM_CLASSES)]),
                   out_tasks=TASKS, return_token_types=(0, 1, 2, 3, 4))


def quad_state_dict(cfg: OracleConfig, seed: int = 0, dim_class_emb: int = 16):
    sd = init_state_dict(cfg, seed=seed)
    g = torch.Generator().manual_seed(seed + 55)
    D, P = cfg.dim, cfg.patch
    del sd["input_adapters.dnw.proj.weight"]
    sd["input_adapters.dnw.class_emb.weight"] = _trunc_normal((NUM_CLASSES, dim_class_emb), 0.02, g)
    sd["input_adapters.dnw.proj.weight"] = _xavier(D, dim_class_emb * P * P, g).reshape(D, dim_class_emb, P, P)
    return sd


def quad_forward(sd, cfg: OracleConfig, x: Dict[str, torch.Tensor], num_encoded_tokens: int, task_masks=None,
                 alphas=1.0, sample_tasks_uniformly: bool = False):
    """-> (preds, task_masks, return_tokens, ori_tokens, encoder_fusion_tokens)"""
    B, _, H, W = x["s1"].shape
    dev = x["s1"].device
    Fn = cfg.num_patches
    tok = OrderedDict()
    for t in x:                                   # dict order of the caller, like the reference (:337-344)
        if t == "dnw":
            tok[t] = semseg_embed(sd, "input_adapters.dnw.", x[t], cfg.patch)
        elif t in TASKS:
            tok[t] = patch_embed(sd, f"input_adapters.{t}.", x[t], cfg)
    fusion = add_fusion_posemb(sd, sd["fusion_tokens"].expand(B, -1, -1))
    n_per_task = OrderedDict((t, v.shape[1]) for t, v in tok.items())
    if task_masks is None:
        task_masks, ids_keep, ids_restore = generate_random_masks(n_per_task, B, num_encoded_tokens, dev, alphas=alphas,
                                                                  sample_tasks_uniformly=sample_tasks_uniformly)
    else:
        ids_keep, ids_restore = masks_from_task_masks(task_masks, list(tok.keys()))
    idx = {t: (task_masks[t][0] == 0).nonzero(as_tuple=True)[0] for t in TASKS}
    counts = [len(idx[t]) for t in TASKS]
    tokens = torch.cat([tok[t][:, idx[t]] for t in TASKS] + [fusion], dim=1)           # (:397-409)
    nenc = num_encoded_tokens
    types = torch.tensor(sum(([TYPE[t]] * n for t, n in zip(TASKS, counts)), []) + [FUSION] * Fn, device=dev)
    zmask = (types[:, None] == types[None, :]) | (types[:, None] == FUSION)            # (:411-431)
    for i in range(cfg.depth):
        tokens = zorro_block(sd, f"blocks.{i}.", tokens, zmask, cfg)
    tokens = zorro_layer_norm(tokens, sd["norm.gamma"], cfg)
    rtypes = torch.tensor(list(cfg.return_token_types), device=dev)
    pmask = (rtypes[:, None] == types[None, :]) | (rtypes[:, None] == FUSION)          # (:449-454)
    r = zorro_attention(sd, "attn_pool.", sd["return_tokens"].expand(B, -1, -1), cfg, context=tokens, attn_mask=pmask)
    return_tokens = r + biased_mlp(sd, "mlp.", zorro_layer_norm(r, sd["norm.gamma"], cfg), cfg)
    enc_fusion = tokens[:, nenc:]
    preds = {t: simple_output_adapter(sd, t, enc_fusion, (H, W), cfg) for t in cfg.out_tasks}
    return preds, task_masks, return_tokens, tokens[:, :nenc], enc_fusion


def quad_loss(out, targets, cfg: OracleConfig):
    """pretrain_mmae_my.py DOMAIN_CONF losses summed over the output domains (same assembly as pretrain_mmae.py:476-487)"""
    preds, masks = out[0], out[1]
    total = 0
    for t, p in preds.items():
        if t == "dnw":
            total = total + masked_ce_loss(p.float(), targets[t], masks[t], cfg.patch)
        elif t == "dem":
            total = total + masked_l1_loss(p.float(), targets[t], masks[t], cfg.patch)
        else:
            total = total + masked_mse_loss(p.float(), targets[t], masks[t], cfg.patch)
    return total
