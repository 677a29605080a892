#!/usr/bin/env python
"""FusedAdamW.step() alone on the cfg-2 model's parameters (gradients = noise, bf16 weight images registered by one
forward): ms per step and algorithmic GB/s (p, g, m, v read; p, m, v written; bf16 image written = 30 B / element)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_batch
from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model
torch.manual_seed(0)
B = int(os.environ.get("AB_BATCH", "32"))
model = build_pretrain_model("base", "crossattn", image_size=224).cuda()
step = PretrainStep(model, num_encoded_tokens=294, global_batch=B)
x = {k: v.cuda() for k, v in synthetic_batch(B, 224, 1234).items()}
for i in range(2):
    step(x)          # registers the weight images, leaves gradients in place
n = sum(p.numel() for p in model.parameters() if p.grad is not None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): step.opt.step()
torch.cuda.synchronize()
e0.record()
for _ in range(20): step.opt.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("adamw: %d elements with gradients, %.3f ms per step, %.0f GB/s of 30 B/element" % (n, ms, n * 30 / ms / 1e6))
