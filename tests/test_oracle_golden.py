"""CPU: the oracle restatement vs. tensors produced by the reference's own code (tests/golden/*.pt,
made by tests/golden/make_golden.py).  This is what pins the oracle."""
import os
from collections import OrderedDict

import pytest
import torch

import oracle
from oracle import OracleConfig


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _cfg(d):
    return OracleConfig(**d)


def _inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((t, torch.randn(batch, C, cfg.image_size, cfg.image_size, generator=g))
                       for t, C in cfg.channels.items())


def _sd(cfg, fx=None, grad=False):
    sd = oracle.perturb_state_dict(oracle.init_state_dict(cfg, seed=0), seed=7)
    if grad:
        for k, v in sd.items():
            if not (k.endswith(".beta") or k.endswith("pos_emb")):
                v.requires_grad_(True)
    return sd


@pytest.mark.parametrize("name", ["crossattn_simple", "plain_xattn", "crossattn_uniform"])
def test_model_forward_backward_matches_reference(golden_dir, name):
    fx = _load(golden_dir, name)
    cfg = _cfg(fx["cfg"])
    sd = _sd(cfg, grad=True)
    # state_dict schema (names, shapes) is the reference's
    assert {k: tuple(v.shape) for k, v in sd.items()} == dict(fx["state_dict_keys"])
    x = _inputs(cfg, fx["batch"], fx["input_seed"])
    torch.manual_seed(fx["mask_seed"])
    out = oracle.multimae_forward(sd, cfg, x, num_encoded_tokens=fx["nenc"], alphas=1.0,
                                  sample_tasks_uniformly=fx["uniformly"])
    for t in fx["task_masks"]:
        assert torch.equal(out[1][t], fx["task_masks"][t])          # int64, bit-exact
    for t in fx["preds"]:
        torch.testing.assert_close(out[0][t], fx["preds"][t], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[2], fx["return_tokens"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[3], fx["ori_tokens"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[4], fx["fusion_tokens"], rtol=1e-5, atol=1e-5)
    for a, b in zip(out[5:], fx["extra_return_tokens"]):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
    loss, _ = oracle.pretrain_loss(out, x, cfg)
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    got = {k for k, v in sd.items() if v.requires_grad and v.grad is not None}
    assert got == set(fx["grad_norms"])                               # same set of params gets a gradient
    for k, n in fx["grad_norms"].items():
        torch.testing.assert_close(sd[k].grad.norm(), n, rtol=2e-4, atol=1e-6)
    for k, g in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, g, rtol=1e-4, atol=2e-6)


def test_modality_subsets_match_reference(golden_dir):
    fx = _load(golden_dir, "subsets")
    cfg = _cfg(fx["cfg"])
    sd = _sd(cfg)
    x = _inputs(cfg, 2, fx["input_seed"])
    Fn = cfg.num_patches
    for key, ref in fx["results"].items():
        present = key.split("+")
        tm = {t: (torch.zeros if t in present else torch.ones)(1, Fn, dtype=torch.long) for t in ("s1", "s2", "dem")}
        with torch.no_grad():
            out = oracle.multimae_forward(sd, cfg, x, task_masks=tm, num_encoded_tokens=Fn * len(present))
        torch.testing.assert_close(out[2], ref["return_tokens"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out[3], ref["ori_tokens"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out[4], ref["fusion_tokens"], rtol=1e-5, atol=1e-5)
        for t in ref["preds"]:
            torch.testing.assert_close(out[0][t], ref["preds"][t], rtol=1e-5, atol=1e-5)
        for a, b in zip(out[5:], ref["extra_return_tokens"]):
            # an absent modality pools over zero keys -> NaN in the reference too
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5, equal_nan=True)


def test_losses_match_reference(golden_dir):
    fx = _load(golden_dir, "losses")
    g = torch.Generator().manual_seed(fx["seed"])
    pred = torch.randn(3, 2, 32, 32, generator=g)
    tgt = torch.randn(3, 2, 32, 32, generator=g)
    mask = (torch.rand(3, 16, generator=g) > 0.5).long()
    mask[2] = 0
    a = torch.randn(6, 48, generator=g)
    b = torch.randn(6, 48, generator=g)
    torch.testing.assert_close(oracle.masked_mse_loss(pred, tgt, mask, 8), fx["mse"])
    torch.testing.assert_close(oracle.masked_l1_loss(pred, tgt, mask, 8), fx["l1"])
    torch.testing.assert_close(oracle.masked_mse_loss(pred, tgt, None, 8), fx["mse_nomask"])
    assert float(oracle.masked_mse_loss(pred, tgt, torch.zeros_like(mask), 8)) == float(fx["mse_zeromask"]) == 0.0
    torch.testing.assert_close(oracle.hard_negative_loss(a, b), fx["hardneg"])
    torch.testing.assert_close(oracle.dino_loss(a, b), fx["dino"])
    # masked cross-entropy over class maps (criterion.py:24-58; the loss of the semantic modality)
    logits = torch.randn(3, 5, 32, 32, generator=g)
    cls = torch.randint(0, 5, (3, 32, 32), generator=g)
    torch.testing.assert_close(oracle.masked_ce_loss(logits, cls, mask, 8), fx["ce"])
    torch.testing.assert_close(oracle.masked_ce_loss(logits, cls, None, 8), fx["ce_nomask"])
    assert float(oracle.masked_ce_loss(logits, cls, torch.zeros_like(mask), 8)) == float(fx["ce_zeromask"]) == 0.0


def test_mask_sampler_bit_exact(golden_dir):
    fx = _load(golden_dir, "masks")
    n = OrderedDict((t, fx["num_patches"]) for t in ("s1", "s2", "dem"))
    for case in fx["cases"]:
        torch.manual_seed(case["seed"])
        tm, keep, restore = oracle.generate_random_masks(n, fx["batch"], fx["nenc"], "cpu", alphas=1.0,
                                                         sample_tasks_uniformly=case["uniformly"])
        for t in n:
            assert torch.equal(tm[t], case["task_masks"][t])
        assert torch.equal(keep, case["ids_keep"])
        assert torch.equal(restore, case["ids_restore"])
        assert sum(int((tm[t][0] == 0).sum()) for t in n) == fx["nenc"]


def test_all_masked_row_is_uniform():
    """Appendix A #4: masked_fill(-finfo.max) + softmax on a row with no allowed key == uniform."""
    cfg = OracleConfig(dim=64, depth=1, heads=1, image_size=16, patch=8)
    sd = oracle.init_state_dict(cfg, 0)
    ctx = torch.randn(2, 5, 64)
    q = torch.randn(1, 2, 64)
    mask = torch.tensor([[True, False, True, False, False], [False] * 5])
    out = oracle.zorro_attention(sd, "attn_pool.", q.expand(2, -1, -1), cfg, context=ctx, attn_mask=mask)
    v = torch.nn.functional.linear(ctx, sd["attn_pool.to_kv.weight"])[..., 64:]
    uni = torch.nn.functional.linear(v.mean(dim=1), sd["attn_pool.to_out.weight"])
    torch.testing.assert_close(out[:, 1], uni, rtol=1e-5, atol=1e-6)


def test_lstm_s2dsm_variant_matches_reference(golden_dir):
    """BASELINE config 1 path (multimae_lstm_s2dsm.MultiMAE + the pretrain_mmae_s2dsm.py loss assembly): oracle vs the
    reference's own run -- schema, int64 masks, every output, the loss and all parameter gradients"""
    from oracle import lstm_variant as L
    fx = _load(golden_dir, "lstm_s2dsm")
    cfg = _cfg(fx["cfg"])
    sd = oracle.perturb_state_dict(L.init_state_dict(cfg, seed=0), seed=7)
    for k, v in sd.items():
        if not (k.endswith(".beta") or k.endswith("pos_emb")):
            v.requires_grad_(True)
    assert {k: tuple(v.shape) for k, v in sd.items()} == dict(fx["state_dict_keys"])
    x = _inputs(cfg, fx["batch"], fx["input_seed"])
    torch.manual_seed(fx["mask_seed"])
    out = L.multimae_lstm_forward(sd, cfg, x, num_encoded_tokens=fx["nenc"], alphas=1.0)
    for t in fx["task_masks"]:
        assert torch.equal(out[1][t], fx["task_masks"][t])
    for t in fx["preds"]:
        torch.testing.assert_close(out[0][t], fx["preds"][t], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[2], fx["return_tokens"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[3], fx["ori_tokens"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out[4], fx["fusion_tokens"], rtol=1e-5, atol=1e-5)
    loss = L.pretrain_loss_s2dsm(out, x, cfg)
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    got = {k for k, v in sd.items() if v.requires_grad and v.grad is not None}
    assert got == set(fx["grad_norms"])
    for k, n in fx["grad_norms"].items():
        torch.testing.assert_close(sd[k].grad.norm(), n, rtol=2e-4, atol=1e-6)
    for k, g in fx["grads"].items():
        torch.testing.assert_close(sd[k].grad, g, rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("name", ["crossattn_simple", "crossattn_uniform"])
def test_reference_bf16_autocast_golden_noise_floor(golden_dir, name):
    """The bf16-tier fixtures are the reference's own code under torch.autocast('cpu', bfloat16) on the same weights,
    inputs and mask seed as the fp32 fixtures.  This pins what 'bf16 parity' can mean: the reference's bf16 run sits
    ~6e-3 (activations) / ~3.5e-2 (worst small gradient) from its own fp32 run; masks do not depend on the dtype."""
    f32, b16 = _load(golden_dir, name), _load(golden_dir, name + "_bf16")
    assert b16["autocast"] == "cpu bf16" and f32["cfg"] == b16["cfg"] and f32["mask_seed"] == b16["mask_seed"]
    for t in f32["task_masks"]:
        assert torch.equal(f32["task_masks"][t], b16["task_masks"][t])
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    for t in f32["preds"]:
        assert b16["preds"][t].dtype == torch.bfloat16            # decoders run inside autocast (no fp32_output_adapters)
        assert rel(b16["preds"][t], f32["preds"][t]) < 1e-2
    assert rel(b16["return_tokens"], f32["return_tokens"]) < 1e-2
    assert abs(float(b16["loss"]) - float(f32["loss"])) < 1e-2 * abs(float(f32["loss"]))
    assert set(b16["grad_norms"]) == set(f32["grad_norms"])
    worst = max(rel(b16["grads"][k], g) for k, g in f32["grads"].items() if float(g.norm()) > 1e-6)
    assert worst < 5e-2, worst


def test_vit_baseline_matches_reference(golden_dir):
    """downstream ViTBaseline: the oracle's pyramid features for the 7 modality subsets against the
    reference class's own outputs; absent modalities have neither tokens nor a modality-attention slot"""
    from oracle.vit_baseline import vit_baseline_flags, vit_baseline_forward, vit_baseline_state_dict
    fx = _load(golden_dir, "vitbaseline")
    cfg = _cfg(fx["cfg"])
    sd = vit_baseline_state_dict(cfg, seed=fx["sd_seed"])
    assert {k: tuple(v.shape) for k, v in sd.items()} == dict(fx["state_dict_keys"])        # the reference class's schema
    assert vit_baseline_flags(cfg.depth) == fx["flags"]
    x = _inputs(cfg, fx["batch"], fx["input_seed"])
    for name, feats in fx["results"].items():
        present = name.split("+")
        with torch.no_grad():
            out = vit_baseline_forward(sd, cfg, OrderedDict((t, x[t]) for t in present), in_domains=present)
        for a, b in zip(out, feats):
            assert a.shape == b.shape
            assert float((a - b).norm() / b.norm()) < 1e-5, name


def test_semseg_adapter_matches_reference(golden_dir):
    """SemSegInputAdapter: the oracle's restatement against the reference adapter's tokens and gradients"""
    fx = _load(golden_dir, "semseg_adapter")
    for name, c in fx.items():
        sd = {"a." + k: v.clone().requires_grad_(k != "pos_emb") for k, v in c["state_dict"].items()}
        tok = oracle.semseg_embed(sd, "a.", c["x"], 8, padding_idx=c["padding_idx"])
        torch.testing.assert_close(tok, c["tokens"], rtol=1e-5, atol=1e-6)
        (tok * c["w"]).sum().backward()
        for k, g in c["grads"].items():
            torch.testing.assert_close(sd["a." + k].grad, g, rtol=1e-4, atol=1e-6)


def test_quadruplet_variant_matches_reference(golden_dir):
    """4-modality model (multimae_quadruplet.py): semantic `dnw` input through SemSegInputAdapter, five-type
    zorro / pool masks, cross-entropy on the dnw decoder -- the oracle against the reference model's own run"""
    from oracle.quadruplet import NUM_CLASSES, quad_config, quad_forward, quad_loss, quad_state_dict
    fx = _load(golden_dir, "quadruplet")
    cfg = quad_config(**fx["cfg_kwargs"])
    sd = oracle.perturb_state_dict(quad_state_dict(cfg, seed=0), seed=7)
    assert {k: tuple(v.shape) for k, v in sd.items() if not k.endswith(".beta")} == \
        {k: s for k, s in fx["state_dict_keys"] if not k.endswith(".beta")}
    for k, v in sd.items():
        if not (k.endswith(".beta") or k.endswith("pos_emb")):
            v.requires_grad_(True)
    g = torch.Generator().manual_seed(fx["input_seed"])
    x = OrderedDict((t, torch.randn(fx["batch"], cfg.channels[t], 32, 32, generator=g)) for t in ("s1", "s2", "dem"))
    x["dnw"] = torch.randint(0, NUM_CLASSES, (fx["batch"], 32, 32), generator=g)
    torch.manual_seed(fx["mask_seed"])
    out = quad_forward(sd, cfg, x, num_encoded_tokens=fx["nenc"])
    for t in fx["task_masks"]:
        assert torch.equal(out[1][t], fx["task_masks"][t])
    for t in fx["preds"]:
        torch.testing.assert_close(out[0][t], fx["preds"][t], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out[2], fx["return_tokens"], rtol=1e-4, atol=1e-5)
    loss = quad_loss(out, x, cfg)
    torch.testing.assert_close(loss, fx["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    got = {k: v.grad for k, v in sd.items() if v.grad is not None}
    assert set(fx["grad_norms"]) <= set(got)
    for k, gr in fx["grads"].items():
        if float(gr.norm()) > 1e-7:
            assert float((got[k] - gr).norm() / gr.norm()) < 1e-4, k
