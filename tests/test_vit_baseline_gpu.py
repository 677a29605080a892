"""GPU: the drop-in downstream backbone `ViTBaseline` against the reference class's golden outputs for the
7 modality subsets, and forward + backward through the pyramid taps against the oracle."""
import os
from collections import OrderedDict

import pytest
import torch

from oracle import OracleConfig

pytestmark = pytest.mark.gpu


def _build(cfg, sd, in_domains):
    from incomplete_multimodal_fusion_b200.multimae.input_adapters import FusionInputAdapter, PatchedInputAdapter
    from incomplete_multimodal_fusion_b200.multimae.multimae_big_imcomplete import ViTBaseline
    ia = OrderedDict((t, PatchedInputAdapter(num_channels=C, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                     for t, C in cfg.channels.items())
    ia["fusion"] = FusionInputAdapter(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
    m = ViTBaseline(pretrained=None, pretrain_size=cfg.image_size, input_adapters=ia, output_adapters=None,
                    in_domains=list(in_domains), dim_tokens=cfg.dim, depth=cfg.depth, dim_head=cfg.dim_head, heads=cfg.heads,
                    ff_mult=cfg.ff_mult, num_fusion_tokens=cfg.num_patches)
    m.load_state_dict(sd, strict=True)      # same keys and shapes as the reference class (checked in the CPU test)
    return m.cuda()


def _inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((t, torch.randn(batch, C, cfg.image_size, cfg.image_size, generator=g)) for t, C in cfg.channels.items())


def rel(a, b):
    return float((a.detach().float() - b.detach().float()).norm() / b.detach().float().norm().clamp_min(1e-12))


def test_vit_baseline_modality_subsets_match_reference_golden(golden_dir):
    from oracle.vit_baseline import vit_baseline_state_dict
    fx = torch.load(os.path.join(golden_dir, "vitbaseline.pt"), weights_only=False)
    cfg = OracleConfig(**fx["cfg"])
    sd = vit_baseline_state_dict(cfg, seed=fx["sd_seed"])
    x = _inputs(cfg, fx["batch"], fx["input_seed"])
    for name, feats in fx["results"].items():
        present = name.split("+")
        model = _build(cfg, sd, present).eval()
        assert model.flags == fx["flags"]
        with torch.no_grad():
            out = model(OrderedDict((t, x[t].cuda()) for t in present))
        for a, b in zip(out, feats):
            assert a.shape == b.shape
            assert rel(a, b.cuda()) < 1e-2, (name, rel(a, b.cuda()))


def test_vit_baseline_backward_through_taps_matches_oracle():
    from oracle.vit_baseline import vit_baseline_forward, vit_baseline_state_dict
    cfg = OracleConfig(variant="crossattn", decoder="simple", dim=128, depth=4, heads=2, dim_head=64, image_size=64, patch=16,
                       dec_dim=64, dec_depth=1, dec_heads=2)
    sd = vit_baseline_state_dict(cfg, seed=3)
    present = ["s1", "dem"]
    x = _inputs(cfg, 3, 5)
    xs = OrderedDict((t, x[t]) for t in present)
    model = _build(cfg, sd, present).eval()          # eval: every token encoded (training mode draws a random subset)
    # gradients are compared through forward_features (the four taps): `forward` adds a MaxPool2d on the last tap, whose
    # argmax flips under bf16 noise and re-routes the gradient (a property of the head, not of the encoder)
    out, nh, nw = model.forward_features(OrderedDict((t, v.cuda()) for t, v in xs.items()))
    w = [torch.randn(f.shape, generator=torch.Generator().manual_seed(9 + i)).cuda() for i, f in enumerate(out)]
    sum((f * wi).sum() for f, wi in zip(out, w)).backward()
    sdo = OrderedDict((k, v.clone().requires_grad_(v.is_floating_point() and not k.endswith("pos_emb"))) for k, v in sd.items())
    ref = vit_baseline_forward(sdo, cfg, xs, in_domains=present, return_taps=True)
    sum((f * wi.cpu()).sum() for f, wi in zip(ref, w)).backward()
    assert (nh, nw) == (cfg.image_size // cfg.patch,) * 2 and len(out) == 4
    for a, b in zip(out, ref):
        assert rel(a, b.cuda()) < 1e-2
    with torch.no_grad():   # and the full forward (pyramid heads) once more against the oracle
        full = model(OrderedDict((t, v.cuda()) for t, v in xs.items()))
        full_ref = vit_baseline_forward(sd, cfg, xs, in_domains=present)
    for a, b in zip(full, full_ref):
        assert a.shape == b.shape and rel(a, b.cuda()) < 1e-2
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    errs = {}
    for k, v in sdo.items():
        if v.grad is None or float(v.grad.norm()) < 1e-8:
            continue
        assert k in grads, k
        errs[k] = rel(grads[k], v.grad.cuda())
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
    print("ViTBaseline gradient errors vs the fp32 oracle, worst:", worst)
    # bf16 operands against an fp32 oracle: same bound as the pre-training model's gradient tests (measured 1.5e-2)
    assert worst[0][1] < 3e-2, worst
    assert len(errs) > 50
    # the absent modality's adapter and every block's parameters that only it would touch get no gradient
    assert model.input_adapters["s2"].proj.weight.grad is None
