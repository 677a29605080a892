"""GPU: the multi-tensor AdamW / grad-norm kernels against torch.optim.AdamW and clip_grad_norm_, and the bf16 weight
images refreshed by the optimiser launch against a fresh cast of the updated parameters."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(768, 512), (70000,), (3, 5, 7), (1,), (256, 1, 16, 16), (131072 + 3,), (64, 64)]
    return [torch.randn(*s, generator=g, device="cuda") for s in shapes]


@pytest.mark.parametrize("max_norm", [None, 0.5])
def test_fused_adamw_matches_torch(max_norm):
    from incomplete_multimodal_fusion_b200.optim import FusedAdamW
    ours = [torch.nn.Parameter(p.clone()) for p in _params()]
    ref = [torch.nn.Parameter(p.clone()) for p in _params()]
    opt = FusedAdamW(ours, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05, max_grad_norm=max_norm, track_grad_norm=True)
    topt = torch.optim.AdamW(ref, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    for step in range(4):
        grads = _params(seed=10 + step)
        for i, (a, b, g) in enumerate(zip(ours, ref, grads)):
            if step == 1 and i == 2:          # a parameter without a gradient on some step is skipped, its step count too
                a.grad = b.grad = None
                continue
            a.grad, b.grad = g.clone(), g.clone()
        total = torch.nn.utils.clip_grad_norm_(ref, max_norm if max_norm else float("inf"))
        if step == 2:
            opt.param_groups[0]["lr"] = topt.param_groups[0]["lr"] = 1e-3      # schedule code assigns per step
        opt.step()
        topt.step()
        assert abs(float(opt.grad_norm) - float(total)) < 1e-4 * float(total)
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), float((a - b).abs().max())
    for m, v, b in zip(opt.exp_avg, opt.exp_avg_sq, ref):
        # (torch forms m with lerp: m + (1-b1)(g - m); a few ulps apart from b1*m + (1-b1)*g)
        assert torch.allclose(m, topt.state[b]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(v, topt.state[b]["exp_avg_sq"], rtol=1e-5, atol=1e-8)


def test_cosine_scheduler_table():
    from incomplete_multimodal_fusion_b200.optim import cosine_scheduler
    t = cosine_scheduler(1e-3, 1e-6, epochs=10, niter_per_ep=7, warmup_epochs=2, start_warmup_value=1e-6)
    assert len(t) == 70 and abs(t[0] - 1e-6) < 1e-12 and abs(t[13] - 1e-3) < 1e-12 and t[14] == pytest.approx(1e-3)
    assert all(t[i] >= t[i + 1] for i in range(14, 69)) and t[-1] < 2e-6


def test_optimizer_refreshes_bf16_weight_images():
    """two pre-training steps with the fused optimiser: every cached bf16 image equals a fresh cast of its parameters,
    the cache is not rebuilt in between, and the run equals one with torch's AdamW + lazy casts"""
    from collections import OrderedDict
    from incomplete_multimodal_fusion_b200 import functions
    from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model
    losses = {}
    for torch_opt in (False, True):
        functions.WEIGHTS.clear()
        torch.manual_seed(0)
        model = build_pretrain_model("tiny", "crossattn", image_size=64, depth=2).cuda()
        step = PretrainStep(model, num_encoded_tokens=24, global_batch=4, torch_optimizer=torch_opt)
        g = torch.Generator().manual_seed(3)
        x = OrderedDict((t, torch.randn(4, c, 64, 64, generator=g).cuda()) for t, c in (("s1", 1), ("s2", 3), ("dem", 1)))
        out = []
        for i in range(3):
            torch.manual_seed(1 + i)
            out.append(float(step(x)))
        losses[torch_opt] = out
        if not torch_opt:
            checked = 0
            for key, (refs, vers, img) in functions.WEIGHTS._store.items():
                srcs = [r() for r in refs]
                if key[0] not in ("w", "cat", "geglu") or any(s is None for s in srcs):
                    continue
                assert vers == tuple((s.data_ptr(), s._version) for s in srcs), key      # stamped current by the optimiser
                row = 0
                for s in srcs:
                    w2 = s.detach().reshape(s.shape[0], -1)
                    assert torch.equal(img[row:row + w2.shape[0], :w2.shape[1]], w2.to(torch.bfloat16)), key
                    row += w2.shape[0]
                checked += 1
            assert checked > 10
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) < 2e-3 * abs(b), (losses[False], losses[True])


def test_fused_adamw_state_dict_round_trip_with_torch():
    """checkpoint layout = torch.optim.AdamW's (the reference saves / restores optimizer.state_dict(),
    utils/checkpoint.py:83,128): our state loads into torch's AdamW and torch's state loads into ours, and both
    continue identically; the single param group carries `lr_scale` for the reference loop's schedule code"""
    from incomplete_multimodal_fusion_b200.optim import FusedAdamW
    ours = [torch.nn.Parameter(p.clone()) for p in _params()]
    ref = [torch.nn.Parameter(p.clone()) for p in _params()]
    opt = FusedAdamW(ours, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    topt = torch.optim.AdamW(ref, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    assert opt.param_groups[0]["lr_scale"] == 1.0
    for step in range(2):
        for a, b, g in zip(ours, ref, _params(seed=20 + step)):
            a.grad, b.grad = g.clone(), g.clone()
        opt.step()
        topt.step()
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(ours)))
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 2.0
    # ours -> a fresh torch AdamW, torch's -> a fresh FusedAdamW (on copies of the current parameters)
    ours2 = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    ref2 = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt2 = FusedAdamW(ours2, lr=1.0)
    opt2.load_state_dict(topt.state_dict())
    topt2 = torch.optim.AdamW(ref2, lr=1.0)
    topt2.load_state_dict(sd)
    assert opt2.param_groups[0]["lr"] == 3e-3 and opt2.param_groups[0]["betas"] == (0.9, 0.95)
    for a, b, g in zip(ours2, ref2, _params(seed=30)):
        a.grad, b.grad = g.clone(), g.clone()
    opt2.step()
    topt2.step()
    for a, b in zip(ours2, ref2):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), float((a - b).abs().max())
    for i, b in enumerate(ref2):
        d = (opt2.exp_avg[i] - topt2.state[b]["exp_avg"]).abs().max()
        assert torch.allclose(opt2.exp_avg[i], topt2.state[b]["exp_avg"], rtol=1e-5, atol=1e-6), (i, float(d), float(topt2.state[b]["exp_avg"].abs().max()))
        assert opt2.steps[i] == 3
