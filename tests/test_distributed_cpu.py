"""CPU, world_size 2, gloo: the bucketed gradient all-reduce of the data-parallel path (the one collective of the
hot path; reference: DDP at pretrain_mmae.py:342-345) averages gradients across ranks, keeps replicas identical and
skips parameters that never receive a gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from incomplete_multimodal_fusion_b200.training import GradAllReduce
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(n)) for n in (1000, 7, 300000, 12)]
    unused = torch.nn.Parameter(torch.randn(5))
    red = GradAllReduce(params + [unused], bucket_mb=1)
    for step in range(2):
        for i, p in enumerate(params):
            p.grad = torch.full_like(p, float(rank + 1 + i + step))
        red.reduce()
        for i, p in enumerate(params):
            expect = sum(r + 1 + i + step for r in range(world)) / world
            assert torch.allclose(p.grad, torch.full_like(p, expect)), (rank, i)
        assert unused.grad is None
    assert len(red.buckets) >= 2          # 1 MB buckets over ~1.2 MB of gradients
    out.put((rank, float(params[2].grad[0])))
    dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get() for _ in range(2))
    assert res[0] == res[1]
