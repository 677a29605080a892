#!/usr/bin/env python
"""BASELINE config 3: incomplete-modality inference -- ViT-B/16 fusion-block encoder forward (eval, no grad) on each of
the 7 non-empty subsets of {s1, s2, dem}, batch 512, absent modalities expressed through explicit task_masks exactly as
the reference does (infer_mmae.py:344-361).  Prints one line per subset: samples/s and the attention
kernel's allowed-pair fraction (skipped key blocks earn no credit).  With --vitbaseline the caller is the downstream
`ViTBaseline` backbone instead (multimae_big_imcomplete.py: absent modalities have no tokens AND no modality-attention slot;
four pyramid feature maps out), which is how the segmentation models consume the encoder."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synthetic_batch  # noqa: E402
from incomplete_multimodal_fusion_b200.training import build_pretrain_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--size", default="base")
ap.add_argument("--image", type=int, default=224)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--vitbaseline", action="store_true")
a = ap.parse_args()
torch.manual_seed(0)
model = build_pretrain_model(a.size, "crossattn", image_size=a.image).cuda().eval()
model.output_adapters = None                    # encoder only: forward returns (tokens, return_tokens, task_masks)
x = {k: v.cuda() for k, v in synthetic_batch(a.batch, a.image, 1234).items()}
F = (a.image // 16) ** 2
if a.vitbaseline:
    from collections import OrderedDict
    from incomplete_multimodal_fusion_b200.multimae.input_adapters import FusionInputAdapter, PatchedInputAdapter
    from incomplete_multimodal_fusion_b200.multimae.multimae_big_imcomplete import ViTBaseline
    from incomplete_multimodal_fusion_b200.training import SIZES
    dim, depth, heads = SIZES[a.size]
    for bits in range(1, 8):
        present = [t for i, t in enumerate(("s1", "s2", "dem")) if bits >> i & 1]
        ia = OrderedDict((t, PatchedInputAdapter(num_channels=c, stride_level=1, patch_size_full=16, image_size=a.image))
                         for t, c in (("s1", 1), ("s2", 3), ("dem", 1)))
        ia["fusion"] = FusionInputAdapter(num_channels=1, stride_level=1, patch_size_full=16, image_size=a.image)
        torch.manual_seed(0)
        vb = ViTBaseline(pretrained=None, pretrain_size=a.image, input_adapters=ia, output_adapters=None, in_domains=present,
                         dim_tokens=dim, depth=depth, dim_head=64, heads=heads, num_fusion_tokens=F).cuda().eval()
        xs = {t: x[t] for t in present}
        with torch.no_grad():
            for _ in range(2):
                vb(xs)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                feats = vb(xs)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(f"ViTBaseline {'+'.join(present):12s} N={F * (len(present) + 1):4d}  {ms:8.2f} ms/forward  {a.batch / ms * 1e3:9.0f} samples/s  "
              f"maps {[tuple(f.shape[1:]) for f in feats]}  finite={all(bool(torch.isfinite(f).all()) for f in feats)}")
        del vb
    sys.exit(0)
for bits in range(1, 8):
    present = [t for i, t in enumerate(("s1", "s2", "dem")) if bits >> i & 1]
    tm = {t: (torch.zeros if t in present else torch.ones)(1, F, dtype=torch.long, device="cuda") for t in ("s1", "s2", "dem")}
    k = len(present)
    with torch.no_grad():
        for _ in range(2):
            model(x, task_masks=tm, num_encoded_tokens=F * k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            out = model(x, task_masks=tm, num_encoded_tokens=F * k)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    N = F * (k + 1)
    allowed = (k * F * F + F * N) / (N * N)
    print(f"{'+'.join(present):12s} N={N:4d}  {ms:8.2f} ms/forward  {a.batch / ms * 1e3:9.0f} samples/s  allowed-pair fraction {allowed:.2f}"
          f"  finite={bool(torch.isfinite(out[0].float()).all())}")
