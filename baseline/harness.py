"""Builds and runs the reference's own `multimae` classes (baseline/_ref, see tools/make_ref.py) for parity tests and
for the reference legs of bench.py.  Every model here is the reference's stock nn.Module: none of this repository's
kernels, modules or oracle restatements are on its path.

Shapes are described by the same small config object the oracle uses (`oracle.OracleConfig`), only as a carrier of
constructor arguments."""
import contextlib
import io
import os
import sys
import warnings
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(ROOT, "tools") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tools"))


def load(which="refmm"):
    import make_ref
    return make_ref.load(which)


def available():
    try:
        return load() is not None
    except Exception:
        return False


@contextlib.contextmanager
def quiet():
    """the reference prints a tensor shape per forward when sample_tasks_uniformly (multimae.py:177) and emits deprecation
    warnings (torch.cuda.amp.autocast, meshgrid)"""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with contextlib.redirect_stdout(io.StringIO()):
            yield


def build_model(cfg, sd=None, device="cpu"):
    """the reference's MultiMAE for `cfg.variant` in {'plain', 'crossattn', 'lstm_s2dsm'} with its own adapters"""
    ref = load()
    Adapter, Fus = ref.input_adapters.PatchedInputAdapter, ref.input_adapters.FusionInputAdapter
    Out = (ref.output_adapters_simple if cfg.decoder == "simple" else ref.output_adapters).SpatialOutputAdapter
    with quiet():
        ia = OrderedDict((t, Adapter(num_channels=C, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                         for t, C in cfg.channels.items())
        ia["fusion"] = Fus(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
        oa = OrderedDict((t, Out(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch, dim_tokens=cfg.dec_dim,
                                 depth=cfg.dec_depth, num_heads=cfg.dec_heads, use_task_queries=True, task=t,
                                 context_tasks=list(cfg.channels), image_size=cfg.image_size, use_xattn=True))
                         for t in cfg.out_tasks)
        mod = {"crossattn": ref.multimae_crossattn, "lstm_s2dsm": ref.multimae_lstm_s2dsm}.get(cfg.variant, ref.multimae)
        T = ref.zorro_utils.TokenTypes
        model = mod.MultiMAE(input_adapters=ia, output_adapters=oa, dim_tokens=cfg.dim, depth=cfg.depth, dim_head=cfg.dim_head,
                             heads=cfg.heads, ff_mult=cfg.ff_mult, num_fusion_tokens=cfg.num_patches,
                             return_token_types=tuple(T(v) for v in cfg.return_token_types), norm_layer=ref.zorro_utils.LayerNorm)
    if sd is not None:
        res = model.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    return model.to(device)


def build_vit_baseline(cfg, sd=None, device="cpu"):
    """the downstream ViTBaseline (multimae_big_imcomplete.py:534-680) over the reference's own adapters"""
    ref = load("refdown")
    mod = ref.multimae_big_imcomplete
    Adapter, Fus = ref.input_adapters.PatchedInputAdapter, ref.input_adapters.FusionInputAdapter
    with quiet():
        ia = OrderedDict((t, Adapter(num_channels=C, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                         for t, C in cfg.channels.items())
        ia["fusion"] = Fus(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
        model = mod.ViTBaseline(pretrained="/nonexistent", pretrain_size=cfg.image_size, input_adapters=ia, output_adapters=None,
                                in_domains=list(cfg.channels), dim_tokens=cfg.dim, depth=cfg.depth, dim_head=cfg.dim_head,
                                heads=cfg.heads, ff_mult=cfg.ff_mult, num_fusion_tokens=cfg.num_patches)
    if sd is not None:
        res = torch.nn.Module.load_state_dict(model, sd, strict=True)      # (the class overrides load_state_dict leniently)
        assert not res.missing_keys and not res.unexpected_keys
    return model.to(device)


def pretrain_loss(out, x, cfg):
    """train_one_epoch's loss assembly (pretrain_mmae.py:476-500) with the reference's own criterion classes"""
    ref = load()
    mse = ref.criterion.MaskedMSELoss(patch_size=cfg.patch, stride=1)
    l1 = ref.criterion.MaskedL1Loss(patch_size=cfg.patch, stride=1)
    preds, masks = out[0], out[1]
    total = 0
    for t in preds:
        total = total + (l1 if t == "dem" else mse)(preds[t].float(), x[t], mask=masks.get(t))
    if len(out) == 8:
        feats = [f.squeeze(1) for f in torch.chunk(out[2], 4, dim=1)]
        toks = [o.squeeze(1) for o in out[5:8]]
        total = total + 0.3 * sum(ref.criterion.dino_loss_func(toks[i], feats[i]) for i in range(3))
    return total


def pretrain_loss_s2dsm(out, x, cfg, hard_negative=True):
    """pretrain_mmae_s2dsm.py:470-492: MSE(s2) + L1(dem) + HardNegtive_loss over the three return tokens (CUDA only: the
    reference's loss calls .cuda(), criterion.py:242; hard_negative=False leaves it out for CPU timing)"""
    ref = load()
    total = ref.criterion.MaskedMSELoss(patch_size=cfg.patch)(out[0]["s2"].float(), x["s2"], mask=out[1]["s2"]) + \
        ref.criterion.MaskedL1Loss(patch_size=cfg.patch)(out[0]["dem"].float(), x["dem"], mask=out[1]["dem"])
    if hard_negative:
        a, b, c = [t.squeeze(1) for t in torch.chunk(out[2], 3, dim=1)]
        hn = ref.criterion.HardNegtive_loss()
        total = total + hn(a, b) + hn(a, c) + hn(b, c)
    return total


def make_optimizer(model, batch):
    """pretrain_mmae.py:115-125 / utils/optim_factory.py:138-176: AdamW, betas (0.9, 0.95), wd 0.05, lr = blr * batch / 256"""
    return torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4 * batch / 256, betas=(0.9, 0.95),
                             weight_decay=0.05)


def train_step(model, opt, x, cfg, nenc, seed, autocast_device=None):
    """one step of train_one_epoch (pretrain_mmae.py:437-517) on the reference model: forward (random masks drawn after
    torch.manual_seed(seed)), losses, backward, AdamW.  autocast_device: None = fp32, 'cuda' / 'cpu' = bf16 autocast"""
    torch.manual_seed(seed)
    opt.zero_grad(set_to_none=True)
    ctx = torch.autocast(autocast_device, dtype=torch.bfloat16) if autocast_device else contextlib.nullcontext()
    with quiet():
        with ctx:
            out = model(x, mask_inputs=True, num_encoded_tokens=nenc, alphas=1.0, sample_tasks_uniformly=True)
        loss = pretrain_loss(out, x, cfg)
        loss.backward()
        opt.step()
    return loss
