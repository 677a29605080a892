#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the REFERENCE's own code.

Run in the authoring container only (it reads /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

The reference package is loaded *in memory* (nothing is copied into the repo) with the two source
fixes SURVEY.md section 0 documents:
  1. pretraining/multimae/zorro_utils.py:255 contains U+FF1A instead of ':' (SyntaxError);
  2. its Block_Fusion (:243-258) calls CrossAttention with the wrong kwargs; the working class is
     downstream/instance_segmentation/modeling/multimae/zorro_utils.py:243-258.
Weights come from ``oracle.init_state_dict`` + ``perturb_state_dict`` (reproducible anywhere) and
are loaded into the reference model with ``strict=True`` -- which also pins the state_dict schema.
Each fixture stores only inputs' seeds and the reference's outputs / loss / gradients.
"""
import importlib.util
import io
import os
import re
import sys
import types
import warnings
from collections import OrderedDict
from functools import partial

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
PKG = os.path.join(REF, "pretraining", "multimae")
DOWN_ZORRO = os.path.join(REF, "downstream", "instance_segmentation", "modeling", "multimae", "zorro_utils.py")

warnings.filterwarnings("ignore")


def _read(path):
    with io.open(path, "r", encoding="utf-8") as f:
        return f.read().replace("\r\n", "\n")


def _class_src(src, name):
    m = re.search(r"^class %s\(.*?(?=^class |\Z)" % name, src, flags=re.S | re.M)
    return m.group(0)


def load_reference():
    """Import the reference ``multimae`` package as ``refmm`` with the two fixes, from memory."""
    if "refmm" in sys.modules:
        return sys.modules["refmm"]
    pkg = types.ModuleType("refmm")
    pkg.__path__ = [PKG]
    sys.modules["refmm"] = pkg

    def load(mod, patch=None):
        src = _read(os.path.join(PKG, mod + ".py"))
        if patch:
            src = patch(src)
        m = types.ModuleType("refmm." + mod)
        m.__package__ = "refmm"
        m.__file__ = os.path.join(PKG, mod + ".py")
        sys.modules["refmm." + mod] = m
        exec(compile(src, m.__file__, "exec"), m.__dict__)
        setattr(pkg, mod, m)
        return m

    def fix_zorro(src):
        src = src.replace("：", ":")
        good = _class_src(_read(DOWN_ZORRO), "Block_Fusion")
        return src.replace(_class_src(src, "Block_Fusion"), good)

    load("multimae_utils")
    load("zorro_utils", fix_zorro)
    load("output_adapter_utils")
    load("input_adapters")
    load("output_adapters")
    load("output_adapters_simple")
    load("criterion")
    # the reference prints a tensor shape every step when sample_tasks_uniformly (multimae.py:177)
    quiet = lambda s: s.replace("print(rand_per_sample_choice.shape)", "pass")
    load("multimae", quiet)
    load("multimae_crossattn", quiet)
    load("multimae_lstm_s2dsm", quiet)
    load("zorro_utils_quadruplet")
    load("multimae_quadruplet", quiet)
    return pkg


def build_reference_model(cfg, sd):
    ref = load_reference()
    Adapter = ref.input_adapters.PatchedInputAdapter
    FusAdapter = ref.input_adapters.FusionInputAdapter
    Out = (ref.output_adapters_simple if cfg.decoder == "simple" else ref.output_adapters).SpatialOutputAdapter
    ia = OrderedDict((t, Adapter(num_channels=C, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                     for t, C in cfg.channels.items())
    ia["fusion"] = FusAdapter(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
    oa = OrderedDict((t, Out(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch,
                             dim_tokens=cfg.dec_dim, depth=cfg.dec_depth, num_heads=cfg.dec_heads,
                             use_task_queries=True, task=t, context_tasks=list(cfg.channels),
                             image_size=cfg.image_size, use_xattn=True))
                     for t in cfg.out_tasks)
    mod = {"crossattn": ref.multimae_crossattn, "lstm_s2dsm": ref.multimae_lstm_s2dsm}.get(cfg.variant, ref.multimae)
    T = ref.zorro_utils.TokenTypes
    model = mod.MultiMAE(input_adapters=ia, output_adapters=oa, dim_tokens=cfg.dim, depth=cfg.depth,
                         dim_head=cfg.dim_head, heads=cfg.heads, ff_mult=cfg.ff_mult,
                         num_fusion_tokens=cfg.num_patches,
                         return_token_types=tuple(T(v) for v in cfg.return_token_types),
                         norm_layer=ref.zorro_utils.LayerNorm)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model


def make_inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((t, torch.randn(batch, C, cfg.image_size, cfg.image_size, generator=g))
                       for t, C in cfg.channels.items())


def reference_loss(ref, out, x, cfg):
    """pretrain_mmae.py:476-500 with the reference's own criterion classes."""
    mse = ref.criterion.MaskedMSELoss(patch_size=cfg.patch, stride=1)
    l1 = ref.criterion.MaskedL1Loss(patch_size=cfg.patch, stride=1)
    preds, masks = out[0], out[1]
    total = 0
    for t in preds:
        total = total + (l1 if t == "dem" else mse)(preds[t].float(), x[t], mask=masks.get(t))
    if len(out) == 8:
        feats = [f.squeeze(1) for f in torch.chunk(out[2], 4, dim=1)]
        toks = [o.squeeze(1) for o in out[5:8]]
        contra = sum(ref.criterion.dino_loss_func(toks[i], feats[i]) for i in range(3))
        total = total + 0.3 * contra
    return total


def model_case(name, cfg, batch, nenc, mask_seed, uniformly, task_masks=None, autocast=False):
    """autocast=True: the same run under torch.autocast('cpu', bfloat16) -- the reference's OWN code in bf16.  (CUDA
    autocast cannot run in the GPU-less authoring container; the CPU policy casts the same Linear / conv / matmul calls
    to bf16 and keeps the fp32 residual stream and LayerNorms, but leaves softmax in bf16 where CUDA autocast upcasts.)"""
    from oracle import init_state_dict
    from oracle.functional import perturb_state_dict
    ref = load_reference()
    sd = perturb_state_dict(init_state_dict(cfg, seed=0), seed=7)
    model = build_reference_model(cfg, sd)
    x = make_inputs(cfg, batch, seed=1234)
    torch.manual_seed(mask_seed)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out = model(x, mask_inputs=True, task_masks=task_masks, num_encoded_tokens=nenc, alphas=1.0,
                    sample_tasks_uniformly=uniformly)
        loss = reference_loss(ref, out, x, cfg)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    fx = {
        "cfg": cfg.__dict__.copy(), "batch": batch, "nenc": nenc, "mask_seed": mask_seed,
        "uniformly": uniformly, "input_seed": 1234, "sd_seed": 0, "perturb_seed": 7, "autocast": "cpu bf16" if autocast else None,
        "task_masks_in": task_masks,
        "preds": {t: v.detach() for t, v in out[0].items()},
        "task_masks": {t: v for t, v in out[1].items()},
        "return_tokens": out[2].detach(), "ori_tokens": out[3].detach(), "fusion_tokens": out[4].detach(),
        "extra_return_tokens": [o.detach() for o in out[5:]],
        "loss": loss.detach(),
        "grad_norms": {k: v.norm() for k, v in grads.items()},
        # a few full gradients (small tensors) + the big ones as norms only keeps the fixture small
        "grads": {k: v for k, v in grads.items() if v.numel() <= 4096 or k.endswith("blocks.0.attn.to_q.weight")},
        "state_dict_keys": [(k, tuple(v.shape)) for k, v in model.state_dict().items()],
        "no_grad_params": [k for k, p in model.named_parameters() if p.requires_grad and p.grad is None],
    }
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "n_grads", len(grads), "counts",
          [int((m[0] == 0).sum()) for m in out[1].values()])


def lstm_case(name, cfg, batch, nenc, mask_seed):
    """BASELINE config 1 path: multimae_lstm_s2dsm.MultiMAE + the loss assembly of pretrain_mmae_s2dsm.py:470-492"""
    from oracle.functional import perturb_state_dict
    from oracle.lstm_variant import init_state_dict as lstm_init
    ref = load_reference()
    sd = perturb_state_dict(lstm_init(cfg, seed=0), seed=7)
    model = build_reference_model(cfg, sd)
    x = make_inputs(cfg, batch, seed=1234)
    torch.manual_seed(mask_seed)
    out = model(x, mask_inputs=True, num_encoded_tokens=nenc, alphas=1.0, sample_tasks_uniformly=False)
    mse = ref.criterion.MaskedMSELoss(patch_size=cfg.patch, stride=1)
    l1 = ref.criterion.MaskedL1Loss(patch_size=cfg.patch, stride=1)
    torch.Tensor.cuda = lambda self, *a, **k: self   # HardNegtive_loss hard-codes .cuda() (criterion.py:242)
    hn = ref.criterion.HardNegtive_loss()
    loss = mse(out[0]["s2"].float(), x["s2"], mask=out[1]["s2"]) + l1(out[0]["dem"].float(), x["dem"], mask=out[1]["dem"])
    a, b, c = [t.squeeze() for t in torch.chunk(out[2], 3, dim=1)]
    loss = loss + hn(a, b) + hn(a, c) + hn(b, c)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    fx = {
        "cfg": cfg.__dict__.copy(), "batch": batch, "nenc": nenc, "mask_seed": mask_seed, "input_seed": 1234,
        "preds": {t: v.detach() for t, v in out[0].items()}, "task_masks": dict(out[1]),
        "return_tokens": out[2].detach(), "ori_tokens": out[3].detach(), "fusion_tokens": out[4].detach(),
        "loss": loss.detach(), "grad_norms": {k: v.norm() for k, v in grads.items()},
        "grads": {k: v for k, v in grads.items() if v.numel() <= 4096 or k.endswith("blocks.0.attn.to_q.weight")
                  or k == "attn_lstm.lstm.weight_hh_l0"},
        "state_dict_keys": [(k, tuple(v.shape)) for k, v in model.state_dict().items()],
        "no_grad_params": [k for k, p in model.named_parameters() if p.requires_grad and p.grad is None],
    }
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "loss", float(loss), "n_grads", len(grads), "counts", [int((m[0] == 0).sum()) for m in out[1].values()])


def subset_case(cfg):
    """The 7 non-empty modality subsets through explicit task_masks (infer_mmae.py:344-361)."""
    from oracle import init_state_dict
    from oracle.functional import perturb_state_dict
    sd = perturb_state_dict(init_state_dict(cfg, seed=0), seed=7)
    model = build_reference_model(cfg, sd).eval()
    x = make_inputs(cfg, 2, seed=99)
    Fn = cfg.num_patches
    res = {}
    for bits in range(1, 8):
        present = [t for i, t in enumerate(("s1", "s2", "dem")) if bits >> i & 1]
        tm = {t: (torch.zeros if t in present else torch.ones)(1, Fn, dtype=torch.long) for t in ("s1", "s2", "dem")}
        with torch.no_grad():
            out = model(x, mask_inputs=True, task_masks=tm, num_encoded_tokens=Fn * len(present))
        res["+".join(present)] = {
            "return_tokens": out[2], "ori_tokens": out[3], "fusion_tokens": out[4],
            "preds": out[0], "extra_return_tokens": list(out[5:]),
        }
    torch.save({"cfg": cfg.__dict__.copy(), "input_seed": 99, "results": res}, os.path.join(HERE, "subsets.pt"))
    print("subsets", list(res))


def loss_case():
    ref = load_reference()
    g = torch.Generator().manual_seed(5)
    pred = torch.randn(3, 2, 32, 32, generator=g)
    tgt = torch.randn(3, 2, 32, 32, generator=g)
    mask = (torch.rand(3, 16, generator=g) > 0.5).long()
    mask[2] = 0                                     # a zero-mask sample -> nanmean path
    a = torch.randn(6, 48, generator=g)
    b = torch.randn(6, 48, generator=g)
    torch.Tensor.cuda = lambda self, *a, **k: self   # HardNegtive_loss hard-codes .cuda() (criterion.py:242)
    fx = {
        "seed": 5,
        "mse": ref.criterion.MaskedMSELoss(patch_size=8)(pred, tgt, mask),
        "l1": ref.criterion.MaskedL1Loss(patch_size=8)(pred, tgt, mask),
        "mse_nomask": ref.criterion.MaskedMSELoss(patch_size=8)(pred, tgt),
        "mse_zeromask": ref.criterion.MaskedMSELoss(patch_size=8)(pred, tgt, torch.zeros_like(mask)),
        "hardneg": ref.criterion.HardNegtive_loss()(a, b),
        "dino": ref.criterion.dino_loss_func(a, b),
    }
    # masked cross-entropy over class maps (criterion.py:24-58): drawn AFTER everything above so the older entries keep
    # their values
    logits = torch.randn(3, 5, 32, 32, generator=g)
    cls = torch.randint(0, 5, (3, 32, 32), generator=g)
    ce = ref.criterion.MaskedCrossEntropyLoss(patch_size=8)
    fx.update({"ce": ce(logits, cls, mask), "ce_nomask": ce(logits, cls),
               "ce_zeromask": ce(logits, cls, torch.zeros_like(mask)).float()})
    torch.save(fx, os.path.join(HERE, "losses.pt"))
    print("losses", {k: float(v) for k, v in fx.items() if k != "seed"})


def mask_case():
    """Mask sampler on the CPU generator, several seeds, both sampling modes (int64, exact)."""
    ref = load_reference()
    from oracle import OracleConfig
    cfg = OracleConfig(dim=64, depth=1, heads=1, image_size=64, patch=8)
    from oracle import init_state_dict
    model = build_reference_model(cfg, init_state_dict(cfg, 0))
    n = OrderedDict((t, torch.zeros(3, cfg.num_patches, 1)) for t in ("s1", "s2", "dem"))
    cases = []
    for seed in range(12):
        for uni in (False, True):
            torch.manual_seed(seed)
            tm, keep, restore = model.generate_random_masks(n, 96, alphas=1.0, sample_tasks_uniformly=uni)
            cases.append({"seed": seed, "uniformly": uni, "task_masks": tm, "ids_keep": keep, "ids_restore": restore})
    torch.save({"num_patches": cfg.num_patches, "nenc": 96, "batch": 3, "cases": cases}, os.path.join(HERE, "masks.pt"))
    print("masks", len(cases))


def semseg_case():
    """SemSegInputAdapter (input_adapters.py:209-328) on its own: tokens and parameter gradients for a random class map,
    with and without a padding class"""
    ref = load_reference()
    out = {}
    for name, pad in (("plain", None), ("padded", 3)):
        torch.manual_seed(21)
        ad = ref.input_adapters.SemSegInputAdapter(num_classes=9, stride_level=1, patch_size_full=8, dim_tokens=64, image_size=32,
                                                   dim_class_emb=16, interpolate_class_emb=False, emb_padding_idx=pad)
        g = torch.Generator().manual_seed(22)
        with torch.no_grad():
            ad.proj.bias.copy_(torch.randn(64, generator=g) * 0.1)
        x = torch.randint(0, ad.num_classes, (3, 32, 32), generator=g)
        w = torch.randn(3, 16, 64, generator=g)
        tok = ad(x)
        (tok * w).sum().backward()
        out[name] = {"state_dict": {k: v.detach().clone() for k, v in ad.state_dict().items()}, "x": x, "w": w, "tokens": tok.detach(),
                     "padding_idx": pad, "grads": {k: p.grad.clone() for k, p in ad.named_parameters() if p.grad is not None}}
    torch.save(out, os.path.join(HERE, "semseg_adapter.pt"))
    print("semseg_adapter", {k: tuple(v["tokens"].shape) for k, v in out.items()}, list(out["plain"]["grads"]))


def quadruplet_case():
    """4-modality model (multimae_quadruplet.py) with the semantic `dnw` input and the masked cross-entropy loss"""
    from oracle.quadruplet import NUM_CLASSES, quad_config, quad_state_dict
    from oracle.functional import perturb_state_dict
    ref = load_reference()
    cfg = quad_config(dim=128, depth=2, heads=2, dim_head=64, image_size=32, patch=8, dec_dim=64, dec_depth=1, dec_heads=2)
    sd = perturb_state_dict(quad_state_dict(cfg, seed=0), seed=7)
    Adapter, Sem, Fus = ref.input_adapters.PatchedInputAdapter, ref.input_adapters.SemSegInputAdapter, ref.input_adapters.FusionInputAdapter
    ia = OrderedDict((t, Adapter(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                     for t in ("s1", "s2", "dem"))
    ia["dnw"] = Sem(num_classes=NUM_CLASSES, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size, dim_class_emb=16,
                    interpolate_class_emb=False)
    ia["fusion"] = Fus(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
    Out = ref.output_adapters_simple.SpatialOutputAdapter
    oa = OrderedDict((t, Out(num_channels=cfg.channels[t], stride_level=1, patch_size_full=cfg.patch, dim_tokens=cfg.dec_dim,
                             depth=cfg.dec_depth, num_heads=cfg.dec_heads, use_task_queries=True, task=t,
                             context_tasks=list(cfg.channels), image_size=cfg.image_size, use_xattn=True)) for t in cfg.out_tasks)
    T = ref.zorro_utils_quadruplet.TokenTypes
    model = ref.multimae_quadruplet.MultiMAE(input_adapters=ia, output_adapters=oa, dim_tokens=cfg.dim, depth=cfg.depth,
                                             dim_head=cfg.dim_head, heads=cfg.heads, ff_mult=cfg.ff_mult,
                                             num_fusion_tokens=cfg.num_patches,
                                             return_token_types=(T.S1, T.S2, T.DEM, T.DNW, T.FUSION),
                                             norm_layer=ref.zorro_utils_quadruplet.LayerNorm)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    g = torch.Generator().manual_seed(1234)
    x = OrderedDict((t, torch.randn(3, cfg.channels[t], 32, 32, generator=g)) for t in ("s1", "s2", "dem"))
    x["dnw"] = torch.randint(0, NUM_CLASSES, (3, 32, 32), generator=g)
    torch.manual_seed(5)
    out = model(x, mask_inputs=True, num_encoded_tokens=28, alphas=1.0, sample_tasks_uniformly=False)
    mse, l1 = ref.criterion.MaskedMSELoss(patch_size=cfg.patch), ref.criterion.MaskedL1Loss(patch_size=cfg.patch)
    ce = ref.criterion.MaskedCrossEntropyLoss(patch_size=cfg.patch)
    preds, masks = out[0], out[1]
    loss = mse(preds["s1"].float(), x["s1"], mask=masks["s1"]) + mse(preds["s2"].float(), x["s2"], mask=masks["s2"]) + \
        l1(preds["dem"].float(), x["dem"], mask=masks["dem"]) + ce(preds["dnw"].float(), x["dnw"], mask=masks["dnw"])
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    torch.save({"cfg_kwargs": dict(dim=128, depth=2, heads=2, dim_head=64, image_size=32, patch=8, dec_dim=64, dec_depth=1, dec_heads=2),
                "batch": 3, "nenc": 28, "mask_seed": 5, "input_seed": 1234,
                "preds": {t: v.detach() for t, v in preds.items()}, "task_masks": dict(masks), "return_tokens": out[2].detach(),
                "ori_tokens": out[3].detach(), "fusion_tokens": out[4].detach(), "loss": loss.detach(),
                "grad_norms": {k: v.norm() for k, v in grads.items()},
                "grads": {k: v for k, v in grads.items() if v.numel() <= 4096 or k.endswith("blocks.0.attn.to_q.weight")
                          or k == "input_adapters.dnw.class_emb.weight"},
                "state_dict_keys": [(k, tuple(v.shape)) for k, v in model.state_dict().items()]},
               os.path.join(HERE, "quadruplet.pt"))
    print("quadruplet loss", float(loss), "counts", [int((m[0] == 0).sum()) for m in masks.values()], "n_grads", len(grads))


def load_reference_downstream():
    """the downstream package's modules the ViTBaseline file needs, loaded in memory as `refdown` (the package __init__
    pulls in detectron2; the four files below are self-contained and unpatched)"""
    if "refdown" in sys.modules:
        return sys.modules["refdown"]
    base = os.path.join(REF, "downstream", "instance_segmentation", "modeling", "multimae")
    pkg = types.ModuleType("refdown")
    pkg.__path__ = [base]
    sys.modules["refdown"] = pkg
    for mod in ("multimae_utils", "zorro_utils", "input_adapters", "multimae_big_imcomplete"):
        m = types.ModuleType("refdown." + mod)
        m.__package__ = "refdown"
        m.__file__ = os.path.join(base, mod + ".py")
        sys.modules["refdown." + mod] = m
        exec(compile(_read(m.__file__), m.__file__, "exec"), m.__dict__)
        setattr(pkg, mod, m)
    return pkg


def vitbaseline_case(cfg):
    """downstream ViTBaseline (multimae_big_imcomplete.py), eval mode, for the 7 non-empty modality subsets"""
    from oracle.vit_baseline import vit_baseline_state_dict
    ref = load_reference_downstream()
    mod = ref.multimae_big_imcomplete
    Adapter, FusAdapter = ref.input_adapters.PatchedInputAdapter, ref.input_adapters.FusionInputAdapter
    ia = OrderedDict((t, Adapter(num_channels=C, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size))
                     for t, C in cfg.channels.items())
    ia["fusion"] = FusAdapter(num_channels=1, stride_level=1, patch_size_full=cfg.patch, image_size=cfg.image_size)
    model = mod.ViTBaseline(pretrained="/nonexistent", pretrain_size=cfg.image_size, input_adapters=ia, output_adapters=None,
                            in_domains=list(cfg.channels), dim_tokens=cfg.dim, depth=cfg.depth, dim_head=cfg.dim_head,
                            heads=cfg.heads, ff_mult=cfg.ff_mult, num_fusion_tokens=cfg.num_patches).eval()
    sd = vit_baseline_state_dict(cfg, seed=0)
    res = torch.nn.Module.load_state_dict(model, sd, strict=True)      # (the class overrides load_state_dict leniently)
    assert not res.missing_keys and not res.unexpected_keys
    x = make_inputs(cfg, 2, seed=77)
    out = {}
    for bits in range(1, 8):
        present = [t for i, t in enumerate(("s1", "s2", "dem")) if bits >> i & 1]
        model.in_domains = present          # eval mode encodes `in_domains`; the caller narrows it to what is available
        with torch.no_grad():
            feats = model({t: x[t] for t in present})
        out["+".join(present)] = [f.clone() for f in feats]
    torch.save({"cfg": cfg.__dict__.copy(), "input_seed": 77, "batch": 2, "sd_seed": 0, "flags": list(model.flags),
                "state_dict_keys": [(k, tuple(v.shape)) for k, v in model.state_dict().items()], "results": out},
               os.path.join(HERE, "vitbaseline.pt"))
    print("vitbaseline", list(out), [tuple(f.shape) for f in out["s1+s2+dem"]])


def main():
    from oracle import OracleConfig
    only = sys.argv[1:]     # e.g. `make_golden.py lstm_s2dsm` regenerates just that fixture
    small = dict(dim=128, depth=2, heads=2, dim_head=64, image_size=32, patch=8, dec_dim=64, dec_depth=1, dec_heads=2)
    if not only or "semseg" in only:
        semseg_case()
    if not only or "quadruplet" in only:
        quadruplet_case()
    if not only or "vitbaseline" in only:
        vitbaseline_case(OracleConfig(variant="crossattn", decoder="simple", dim=64, depth=4, heads=1, dim_head=64, image_size=32,
                                      patch=8, dec_dim=64, dec_depth=1, dec_heads=2))
    if not only or "bf16" in only:
        model_case("crossattn_simple_bf16", OracleConfig(variant="crossattn", decoder="simple", **small), 2, 24, 1, False, autocast=True)
        model_case("crossattn_uniform_bf16", OracleConfig(variant="crossattn", decoder="simple", **small), 3, 20, 11, True, autocast=True)
    if not only:
        model_case("crossattn_simple", OracleConfig(variant="crossattn", decoder="simple", **small), 2, 24, 1, False)
        model_case("plain_xattn", OracleConfig(variant="plain", decoder="xattn", **small), 2, 24, 3, True)
        model_case("crossattn_uniform", OracleConfig(variant="crossattn", decoder="simple", **small), 3, 20, 11, True)
    from oracle.lstm_variant import lstm_config
    if not only or "lstm_s2dsm" in only:
        lstm_case("lstm_s2dsm", lstm_config(decoder="simple", **small), 3, 12, 2)
    if only:
        return
    subset_case(OracleConfig(variant="crossattn", decoder="simple", **small))
    loss_case()
    mask_case()


if __name__ == "__main__":
    main()
