#!/usr/bin/env python
"""Run a few pre-training steps of the bench workload (for ncu): python tools/profile_step.py --steps 2 --batch 256"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synthetic_batch  # noqa: E402
from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", default="base")
ap.add_argument("--variant", default="crossattn")
ap.add_argument("--nenc", type=int, default=294)
ap.add_argument("--image", type=int, default=224)
a = ap.parse_args()
torch.manual_seed(0)
model = build_pretrain_model(a.size, a.variant, image_size=a.image).cuda()
step = PretrainStep(model, num_encoded_tokens=a.nenc, global_batch=a.batch)
x = {k: v.cuda() for k, v in synthetic_batch(a.batch, a.image, 1234).items()}
for i in range(a.steps):
    torch.manual_seed(1 + i)
    if i == a.steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()   # ncu --profile-from-start off: every thread (autograd too) of the last step
    loss = step(x)
    if i == a.steps - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
torch.cuda.synchronize()
print("loss", float(loss))
