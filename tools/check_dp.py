#!/usr/bin/env python
"""Data-parallel consistency check on N GPUs (torchrun): one pre-training step with the overlapped reducer (per-layer
in-place all-reduce from inside the encoder backward + finish()) must leave on every rank the SUM over ranks of the
gradients a rank computes alone -- compared against plain dist.all_reduce of the local gradients of an identical second
run with the hooks off.   torchrun --nproc-per-node 2 tools/check_dp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import synthetic_batch  # noqa: E402
from incomplete_multimodal_fusion_b200.training import PretrainStep, build_pretrain_model  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
model = build_pretrain_model("tiny", "crossattn", image_size=64, depth=3).cuda()
for p in model.parameters():
    dist.broadcast(p.data, 0)
x = {k: v.cuda() for k, v in synthetic_batch(8, 64, 100 + rank).items()}


def grads(with_hooks):
    step = PretrainStep(model, num_encoded_tokens=24, global_batch=8 * world)
    if not with_hooks:
        model.grad_hook = model.grad_hook_inplace = None
    model.zero_grad(set_to_none=True)
    torch.manual_seed(5)
    out = model(x, num_encoded_tokens=24, sample_tasks_uniformly=True)
    (step.loss(out, x) / world).backward()
    if with_hooks:
        step.reducer.finish()
        torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


a = grads(True)
b = grads(False)
for v in b.values():
    dist.all_reduce(v)
assert set(a) == set(b)
worst = max(float((a[k] - b[k]).abs().max() / b[k].abs().max().clamp_min(1e-20)) for k in a)
same = all(torch.equal(a[k], t) for k in a for t in [a[k].clone()])
gathered = [torch.zeros(1, device="cuda") for _ in range(world)]
chk = torch.stack([v.double().sum() for v in a.values()]).sum().float().view(1)
dist.all_gather(gathered, chk)
if rank == 0:
    print("ranks", world, "params", len(a), "worst relative difference hook-path vs plain all-reduce: %.2e" % worst,
          "| identical checksum on all ranks:", all(float(g) == float(gathered[0]) for g in gathered))
    assert worst < 1e-3, worst     # split-K / atomics order noise only
dist.destroy_process_group()
