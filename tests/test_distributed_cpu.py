"""CPU, world_size 2, gloo: the gradient all-reduce of the data-parallel path (the one collective of the hot path;
reference: DDP at pretrain_mmae.py:342-345).  `reduce_now` is the in-backward per-layer hook, `finish` reduces the
rest; gradients are summed (the 1/world factor is applied to the loss), replicas stay identical, parameters that never
receive a gradient are skipped, and the flat buckets are reused from step to step."""
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
    red = GradAllReduce(params + [unused])
    assert red.enabled()
    pool_ptrs = None
    for step in range(3):
        # "layer" gradients go through the hook while backward would still be running ...
        layer = [torch.full((3, 4), float(rank + step)), torch.full((5,), float(10 * rank + step))]
        got = red.reduce_now(layer)
        assert torch.allclose(got[0], torch.full((3, 4), float(sum(r + step for r in range(world)))))
        assert torch.allclose(got[1], torch.full((5,), float(sum(10 * r + step for r in range(world)))))
        # ... or, when they already sit in one contiguous buffer (the backward's per-layer arena), in place
        flat = torch.zeros(40)
        va, vb = flat[:12].view(3, 4), flat[16:21]
        va.fill_(float(rank + 1)); vb.fill_(float(2 * rank + step))
        red.reduce_inplace(flat[:24], [va, vb])
        assert torch.allclose(va, torch.full((3, 4), float(sum(r + 1 for r in range(world)))))
        assert torch.allclose(vb, torch.full((5,), float(sum(2 * r + step for r in range(world)))))
        assert float(flat[24:].abs().sum()) == 0.0 and va.data_ptr() in red._reduced
        params[1].grad = got[1][:7] if got[1].numel() >= 7 else torch.full_like(params[1], 1.0 + rank)
        # ... the remaining parameters at the end
        for i in (0, 2, 3):
            params[i].grad = torch.full_like(params[i], float(rank + 1 + i + step))
        before1 = params[1].grad.clone()
        red.finish()
        for i in (0, 2, 3):
            expect = float(sum(r + 1 + i + step for r in range(world)))
            assert torch.allclose(params[i].grad, torch.full_like(params[i], expect)), (rank, i)
        if got[1].numel() < 7:
            assert torch.allclose(params[1].grad, torch.full_like(params[1], float(sum(1.0 + r for r in range(world)))))
        else:
            assert torch.equal(params[1].grad, before1)          # reduced by the hook already: not reduced twice
        assert unused.grad is None
        ptrs = [b.data_ptr() for b in red._pool]
        assert pool_ptrs is None or ptrs == pool_ptrs             # persistent buckets
        pool_ptrs = ptrs
    assert len(red._pool) == 2
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
