#!/usr/bin/env python
"""In-step kernel timeline of the bench workload through torch.profiler (CUPTI): per-kernel totals as they run inside a
warm step (not ncu's serialised cold-cache replays), the GPU idle time between kernels and the largest gaps.

  python tools/timeline_step.py --batch 256 [--out gpurun_out/timeline.txt]"""
import argparse
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import synthetic_batch  # noqa: E402
from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--size", default="base")
ap.add_argument("--variant", default="crossattn")
ap.add_argument("--nenc", type=int, default=294)
ap.add_argument("--image", type=int, default=224)
ap.add_argument("--out", default="")
a = ap.parse_args()
torch.manual_seed(0)
model = build_pretrain_model(a.size, a.variant, image_size=a.image).cuda()
step = PretrainStep(model, num_encoded_tokens=a.nenc, global_batch=a.batch)
x = {k: v.cuda() for k, v in synthetic_batch(a.batch, a.image, 1234).items()}
for i in range(4):
    torch.manual_seed(1 + i)
    step(x)
torch.cuda.synchronize()
# three steps back to back, the MIDDLE one analysed (from the end of the first step's AdamW launch to the end of the
# second's): the host is a step ahead there, as in a training loop; a step profiled right after a synchronize shows
# ~1.3 ms of host-bound gaps in its first 2 ms (mask draws, im2col) that a running loop does not have
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        torch.manual_seed(9 + i)
        step(x)
    torch.cuda.synchronize()

evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
        evs.append((e.time_range.start, e.time_range.end, e.name))
evs.sort()
ends = [e for s_, e, n in evs if "adamw_kernel" in n]
if len(ends) >= 2:
    evs = [ev for ev in evs if ends[0] <= ev[0] and ev[1] <= ends[1]]
tot = defaultdict(float)
cnt = defaultdict(int)
for s, e, n in evs:
    key = n.split("<")[0].split("(")[0][:60]
    tot[key] += (e - s) / 1e3
    cnt[key] += 1
span = (evs[-1][1] - evs[0][0]) / 1e3
busy_end = evs[0][0]
idle = 0.0
gaps = []
prev = ""
for s, e, n in evs:
    if s > busy_end:
        idle += (s - busy_end) / 1e3
        gaps.append(((s - busy_end) / 1e3, n[:70], prev[:40], (s - evs[0][0]) / 1e3))
    if e >= busy_end:
        prev = n
    busy_end = max(busy_end, e)
lines = ["span %.2f ms, sum of kernel time %.2f ms, idle %.2f ms, %d device activities" % (span, sum(tot.values()), idle, len(evs))]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
    lines.append("%9.3f ms %5.1f%% n=%4d  %s" % (v, 100 * v / span, cnt[k], k))
lines.append("largest gaps (ms, kernel that followed):")
ap_top = int(os.environ.get("MMF_TIMELINE_GAPS", "15"))
for g, n, pv, at in sorted(gaps, reverse=True)[:ap_top]:
    lines.append("   %.3f  at %7.2f ms  %s   <- after %s" % (g, at, n, pv))
small = sum(g[0] for g in gaps if g[0] < 0.02)
lines.append("gaps < 20 us: %.2f ms in %d gaps" % (small, sum(1 for g in gaps if g[0] < 0.02)))
txt = "\n".join(lines)
print(txt)
if a.out:
    open(a.out, "w").write(txt + "\n")
