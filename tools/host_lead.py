#!/usr/bin/env python
"""How far ahead of the GPU does the host run?  Per step: host time to enqueue the whole step (no synchronisation but the
mask-count read-back inside it) against the device time of the step (CUDA events)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import synthetic_batch  # noqa: E402
from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model  # noqa: E402

torch.manual_seed(0)
model = build_pretrain_model("base", "crossattn", image_size=224).cuda()
step = PretrainStep(model, num_encoded_tokens=294, global_batch=256)
x = {k: v.cuda() for k, v in synthetic_batch(256, 224, 1234).items()}
for i in range(4):
    torch.manual_seed(1 + i)
    step(x)
torch.cuda.synchronize()
n = 8
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host = []
ev[0].record()
t_all = time.perf_counter()
for i in range(n):
    torch.manual_seed(100 + i)
    t0 = time.perf_counter()
    step(x)
    host.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
t_enq = (time.perf_counter() - t_all) * 1e3
torch.cuda.synchronize()
dev = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print("host enqueue ms per step:", " ".join("%.1f" % h for h in host))
print("device ms per step:      ", " ".join("%.1f" % d for d in dev))
print("host total %.1f ms for %d steps (device %.1f ms)" % (t_enq, n, sum(dev)))
