"""GPU: the drop-in 4-modality model (multimae_quadruplet.py) against the reference model's golden run and
the oracle: bit-exact sampled masks, outputs / loss <= 1e-2, parameter gradients <= the bf16 gradient bound."""
import os
from collections import OrderedDict

import pytest
import torch

import oracle
from oracle.quadruplet import NUM_CLASSES, quad_config, quad_forward, quad_loss, quad_state_dict

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _build(cfg, sd):
    from incomplete_multimodal_fusion_b200.multimae.input_adapters import FusionInputAdapter, PatchedInputAdapter, SemSegInputAdapter
    from incomplete_multimodal_fusion_b200.multimae.multimae_quadruplet import MultiMAE
    from incomplete_multimodal_fusion_b200.multimae.output_adapters_simple import SpatialOutputAdapter
    from incomplete_multimodal_fusion_b200.multimae.zorro_utils_quadruplet import TokenTypes as T
    ia = OrderedDict((t, PatchedInputAdapter(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                     for t in ("s1", "s2", "dem"))
    ia["dnw"] = SemSegInputAdapter(num_classes=NUM_CLASSES, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size,
                                   dim_class_emb=16, interpolate_class_emb=False)
    ia["fusion"] = FusionInputAdapter(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
    oa = OrderedDict((t, SpatialOutputAdapter(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch,
                                              dim_tokens=cfg.dec_dim, depth=cfg.dec_depth, num_heads=cfg.dec_heads, task=t,
                                              context_tasks=list(cfg.channels), image_size=cfg.image_size)) for t in cfg.out_tasks)
    model = MultiMAE(ia, oa, dim_tokens=cfg.dim, depth=cfg.depth, dim_head=cfg.dim_head, heads=cfg.heads, ff_mult=cfg.ff_mult,
                     num_fusion_tokens=cfg.num_patches, return_token_types=(T.S1, T.S2, T.DEM, T.DNW, T.FUSION))
    model.load_state_dict(sd, strict=True)
    return model.cuda()


def _loss(out, x, patch):
    from incomplete_multimodal_fusion_b200.multimae.criterion import MaskedCrossEntropyLoss, MaskedL1Loss, MaskedMSELoss
    mse, l1, ce = MaskedMSELoss(patch_size=patch), MaskedL1Loss(patch_size=patch), MaskedCrossEntropyLoss(patch_size=patch)
    p, m = out[0], out[1]
    return mse(p["s1"], x["s1"], mask=m["s1"]) + mse(p["s2"], x["s2"], mask=m["s2"]) + l1(p["dem"], x["dem"], mask=m["dem"]) + \
        ce(p["dnw"], x["dnw"], mask=m["dnw"])


def test_quadruplet_matches_reference_golden_and_oracle(golden_dir):
    fx = torch.load(os.path.join(golden_dir, "quadruplet.pt"), weights_only=False)
    cfg = quad_config(**fx["cfg_kwargs"])
    sd = oracle.perturb_state_dict(quad_state_dict(cfg, seed=0), seed=7)
    model = _build(cfg, sd)
    g = torch.Generator().manual_seed(fx["input_seed"])
    x = OrderedDict((t, torch.randn(fx["batch"], cfg.channels[t], 32, 32, generator=g).cuda()) for t in ("s1", "s2", "dem"))
    x["dnw"] = torch.randint(0, NUM_CLASSES, (fx["batch"], 32, 32), generator=g).cuda()
    # the reference's masks passed in explicitly (its run drew them from the CPU generator)
    tm = {t: m.cuda() for t, m in fx["task_masks"].items()}
    out = model(x, task_masks=tm, num_encoded_tokens=fx["nenc"])
    for t in fx["preds"]:
        assert rel(out[0][t], fx["preds"][t].cuda()) < 1e-2, (t, rel(out[0][t], fx["preds"][t].cuda()))
    assert rel(out[2], fx["return_tokens"].cuda()) < 1e-2 and rel(out[4], fx["fusion_tokens"].cuda()) < 1e-2
    loss = _loss(out, x, cfg.patch)
    assert abs(float(loss) - float(fx["loss"])) < 1e-2 * abs(float(fx["loss"]))
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(fx["grad_norms"]) <= set(grads)
    gerr = {k: rel(grads[k], g_.cuda()) for k, g_ in fx["grads"].items() if float(g_.norm()) > 1e-6}
    assert max(gerr.values()) < 5e-2, sorted(gerr.items(), key=lambda kv: -kv[1])[:5]     # GOLDEN_GRAD_TOL of test_model_gpu
    # sampled masks on the device: bit-exact against the oracle drawing from the same generator state
    torch.manual_seed(11)
    out2 = model(x, num_encoded_tokens=fx["nenc"], sample_tasks_uniformly=True)
    sdo = OrderedDict((k, v.cuda()) for k, v in sd.items())
    torch.manual_seed(11)
    with torch.no_grad():
        ref2 = quad_forward(sdo, cfg, x, num_encoded_tokens=fx["nenc"], sample_tasks_uniformly=True)
    for t in ref2[1]:
        assert torch.equal(out2[1][t], ref2[1][t]), t
    for t in ref2[0]:
        assert rel(out2[0][t], ref2[0][t]) < 1e-2, t
