#!/usr/bin/env python
"""Input pipeline kernels (mmf_raster_prep) at the cfg-2 shape: batch 256 of 512 x 512 rasters -> 224 x 224 crops of the
256 x 256 resized image.  CUDA events; bytes = source bytes under the crop window (whole source for the per-image
standardisation) + fp32 output; also the pinned-host copy the raw batch needs against the fp32 tensors it replaces."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from incomplete_multimodal_fusion_b200.utils import multimodal_dfc2023 as D  # noqa: E402

B, S, CR = 256, 512, 224


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rgb = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, device="cuda")
sar = torch.rand(B, 1, S, S, device="cuda") + 0.01
dsm = torch.rand(B, 1, S, S, device="cuda") * 30
np.random.seed(0)
top, left = D.RandomCrop(CR).draw(B)
crop = D.upload_crop((top, left, (CR, CR)), B, "cuda")
f = S // 256
for name, fn, src, whole in (("rgb u8 zscore", D.prepare_rgb, rgb, False), ("sar f32 dB+zscore", D.prepare_sar, sar, False),
                             ("dsm f32 standardise", D.prepare_dsm, dsm, True)):
    C = src.shape[1]
    rd = src.numel() * src.element_size() if whole else B * C * (CR * f) ** 2 * src.element_size()
    wr = B * C * CR * CR * 4
    ms = t(lambda: fn(src, crop))
    print(f"{name:22s} {ms:7.3f} ms  {(rd + wr) / ms / 1e6:7.0f} GB/s  (read {rd / 1e6:.0f} MB, write {wr / 1e6:.0f} MB)")
raw_bytes = sum(x.numel() * x.element_size() for x in (rgb, sar, dsm))
print(f"H2D per batch: raw {raw_bytes / 1e6:.0f} MB (uint8 optical) vs {B * 5 * 256 * 256 * 4 / 1e6:.0f} MB of normalised fp32 256x256 tensors")
