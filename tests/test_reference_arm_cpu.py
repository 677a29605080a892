"""CPU: the reference arm (baseline/_ref, built by tools/make_ref.py from /root/reference) IS the code that produced the
committed golden fixtures: its stock modules, run here on the fixtures' weights / inputs / mask seeds, reproduce the
fixtures' outputs, loss and gradients.  This ties every GPU test that compares against baseline/_ref to the same pinned
reference behaviour as the oracle's golden tests.  Skipped when baseline/_ref is absent (no /root/reference and no earlier
build); `python tools/make_ref.py` creates it."""
import json
import os

import pytest
import torch

import oracle
from oracle import OracleConfig
from _util import default_sd, make_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _harness():
    from baseline import harness as H
    if not H.available():
        pytest.skip("baseline/_ref not built (tools/make_ref.py needs /root/reference)")
    return H


def test_manifest_lists_the_two_documented_edits_only():
    H = _harness()
    man = json.load(open(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json")))
    assert len(man["edits"]) == 3 and any("U+FF1A" in e for e in man["edits"]) and any("Block_Fusion" in e for e in man["edits"])
    assert all(len(v["sha256"]) == 64 for v in man["files"].values()) and len(man["files"]) >= 18
    assert H.load() is not None and H.load("refdown") is not None


@pytest.mark.parametrize("name", ["crossattn_simple", "plain_xattn", "crossattn_uniform"])
def test_reference_arm_reproduces_the_golden_fixtures(golden_dir, name):
    H = _harness()
    fx = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    cfg = OracleConfig(**fx["cfg"])
    model = H.build_model(cfg, default_sd(cfg), "cpu")
    x = make_inputs(cfg, fx["batch"], fx["input_seed"])
    torch.manual_seed(fx["mask_seed"])
    with H.quiet():
        out = model(x, mask_inputs=True, task_masks=fx["task_masks_in"], num_encoded_tokens=fx["nenc"], alphas=1.0,
                    sample_tasks_uniformly=fx["uniformly"])
        loss = H.pretrain_loss(out, x, cfg)
        loss.backward()
    for t in fx["task_masks"]:
        assert torch.equal(out[1][t], fx["task_masks"][t])
    for t in fx["preds"]:
        torch.testing.assert_close(out[0][t], fx["preds"][t], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out[2], fx["return_tokens"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss.detach(), fx["loss"], rtol=1e-6, atol=1e-7)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    for k, g in fx["grads"].items():
        torch.testing.assert_close(grads[k], g, rtol=1e-4, atol=1e-6)
    assert sorted(k for k, p in model.named_parameters() if p.requires_grad and p.grad is None) == sorted(fx["no_grad_params"])
