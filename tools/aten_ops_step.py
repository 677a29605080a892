#!/usr/bin/env python
"""Which torch (aten::) operators still launch kernels inside a training step, with call counts and device time: the glue
around the C-ABI kernels (loss assembly, dtype / layout conversions, gradient accumulation by autograd)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import synthetic_batch
from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model
torch.manual_seed(0)
model = build_pretrain_model("base", "crossattn", image_size=224).cuda()
step = PretrainStep(model, num_encoded_tokens=294, global_batch=256)
x = {k: v.cuda() for k, v in synthetic_batch(256, 224, 1234).items()}
for i in range(3): step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(x)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key.startswith("aten::") and e.self_device_time_total > 0:
        rows.append((e.self_device_time_total / 1e3, e.count, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
print("aten ops with device time in one step: %.3f ms in %d calls" % (sum(r[0] for r in rows), sum(r[1] for r in rows)))
for ms, n, k, sh in rows[:40]:
    print("%8.3f ms  n=%3d  %-28s %s" % (ms, n, k, sh))
