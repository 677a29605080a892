"""helpers shared by the tests: build the drop-in model for an OracleConfig, synthetic inputs"""
from collections import OrderedDict

import torch

import oracle


def make_inputs(cfg, batch, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((t, torch.randn(batch, C, cfg.image_size, cfg.image_size, generator=g).to(device))
                       for t, C in cfg.channels.items())


def build_model(cfg, sd=None, device="cuda"):
    from incomplete_multimodal_fusion_b200.multimae import multimae as m_plain
    from incomplete_multimodal_fusion_b200.multimae import multimae_crossattn as m_cross
    from incomplete_multimodal_fusion_b200.multimae.input_adapters import FusionInputAdapter, PatchedInputAdapter
    if cfg.decoder == "simple":
        from incomplete_multimodal_fusion_b200.multimae.output_adapters_simple import SpatialOutputAdapter
    else:
        from incomplete_multimodal_fusion_b200.multimae.output_adapters import SpatialOutputAdapter
    ia = OrderedDict((t, PatchedInputAdapter(num_channels=C, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                     for t, C in cfg.channels.items())
    ia["fusion"] = FusionInputAdapter(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
    oa = OrderedDict((t, SpatialOutputAdapter(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch,
                                              dim_tokens=cfg.dec_dim, depth=cfg.dec_depth, num_heads=cfg.dec_heads, task=t,
                                              context_tasks=list(cfg.channels), image_size=cfg.image_size))
                     for t in cfg.out_tasks)
    from incomplete_multimodal_fusion_b200.multimae import multimae_lstm_s2dsm as m_lstm
    from incomplete_multimodal_fusion_b200.multimae.zorro_utils import TokenTypes
    mod = {"crossattn": m_cross, "lstm_s2dsm": m_lstm}.get(cfg.variant, m_plain)
    model = mod.MultiMAE(ia, oa, dim_tokens=cfg.dim, depth=cfg.depth, dim_head=cfg.dim_head, heads=cfg.heads,
                         ff_mult=cfg.ff_mult, num_fusion_tokens=cfg.num_patches,
                         return_token_types=tuple(TokenTypes(v) for v in cfg.return_token_types))
    if sd is not None:
        model.load_state_dict(sd, strict=True)
    return model.to(device)


def default_sd(cfg):
    if cfg.variant == "lstm_s2dsm":
        from oracle.lstm_variant import init_state_dict as lstm_init
        return oracle.perturb_state_dict(lstm_init(cfg, seed=0), seed=7)
    return oracle.perturb_state_dict(oracle.init_state_dict(cfg, seed=0), seed=7)


def pretrain_loss_s2dsm_ours(out, targets, patch):
    """pretrain_mmae_s2dsm.py:470-492 with the drop-in criterion classes"""
    from incomplete_multimodal_fusion_b200.multimae.criterion import HardNegtive_loss, MaskedL1Loss, MaskedMSELoss
    total = MaskedMSELoss(patch_size=patch)(out[0]["s2"], targets["s2"], mask=out[1]["s2"]) + \
        MaskedL1Loss(patch_size=patch)(out[0]["dem"], targets["dem"], mask=out[1]["dem"])
    a, b, c = [t.squeeze(1) for t in torch.chunk(out[2], 3, dim=1)]
    hn = HardNegtive_loss()
    return total + hn(a, b) + hn(a, c) + hn(b, c)


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def pretrain_loss_ours(out, targets, patch):
    """the loss assembly of train_one_epoch (pretrain_mmae.py:476-500) with the drop-in criterion classes"""
    from incomplete_multimodal_fusion_b200.multimae.criterion import MaskedL1Loss, MaskedMSELoss, dino_loss_func
    mse, l1 = MaskedMSELoss(patch_size=patch), MaskedL1Loss(patch_size=patch)
    preds, masks = out[0], out[1]
    total = 0
    for t, p in preds.items():
        total = total + (l1 if t == "dem" else mse)(p, targets[t], mask=masks.get(t))
    if len(out) == 8:
        feats = [f.squeeze(1) for f in torch.chunk(out[2], 4, dim=1)]
        toks = [o.squeeze(1) for o in out[5:8]]
        total = total + 0.3 * sum(dino_loss_func(toks[i], feats[i]) for i in range(3))
    return total
