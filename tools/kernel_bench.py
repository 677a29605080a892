#!/usr/bin/env python
"""Micro-benchmark of the non-GEMM kernels at the ViT-B cfg-2 shapes (B=256, nenc=294, F=196, D=768, H=8)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from incomplete_multimodal_fusion_b200 import _lib, kernels as K  # noqa: E402

if os.environ.get("MMF_LIB"):   # e.g. scratch/old_libmmf.so built from another revision, for A/B runs on the same box
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])

bf16, f32 = torch.bfloat16, torch.float32
B, nenc, Fn, D, H = 256, 294, 196, 768, 8
N = nenc + Fn
Mh, Mf = B * nenc, B * Fn
Mt = Mh + Mf
HD = H * 64
dev = "cuda"
only = sys.argv[1:] or None


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


def rnd(*s, dt=bf16):
    return (torch.randn(*s, device=dev)).to(dt)


def report(name, ms, gbytes=None, gflop=None):
    s = f"{name:28s} {ms:8.3f} ms"
    if gbytes:
        s += f"  {gbytes / ms * 1e3:8.0f} GB/s"
    if gflop:
        s += f"  {gflop / ms:8.1f} TFLOP/s"
    print(s)


def want(k):
    return only is None or any(o in k for o in only)


if want("ln"):
    x = rnd(Mt, D, dt=f32); g1 = torch.ones(D, device=dev); g2 = torch.ones(D, device=dev)
    y = torch.empty(Mt, D, dtype=bf16, device=dev); st = torch.empty(Mt, 4, device=dev)
    report("ln2 fwd", t(lambda: K.layernorm_fwd(x, g1, y, g2=g2, stats=st)), Mt * D * 6 / 1e9)
    dy = rnd(Mt, D); dres = rnd(Mt, D, dt=f32); dx = torch.empty(Mt, D, device=dev); dxb = torch.empty(Mt, D, dtype=bf16, device=dev)
    dg1 = torch.zeros(D, device=dev); dg2 = torch.zeros(D, device=dev)
    report("ln2 bwd (+dres,+bf16)", t(lambda: K.layernorm_bwd(dy, x, g1, st, dx, dg1, g2=g2, dres=dres, dx_bf16=dxb, dg2=dg2)), Mt * D * 16 / 1e9)
if want("attn"):
    counts = (98, 98, 98)
    seg = torch.tensor([0, 98, 196, 294, 490], dtype=torch.int32, device=dev)
    A_pairs = sum(c * c for c in counts) + Fn * N
    qkv = rnd(Mt, 3 * HD); o = torch.empty(Mt, HD, dtype=bf16, device=dev); lse = torch.empty(B, H, N, device=dev)
    kw = dict(B=B, H=H, Nq=N, Nk=N, dh=64, scale=0.125, n_head_q=nenc, n_head_k=nenc, seg=seg, nseg=4)
    fl = 4 * 64 * A_pairs * B * H / 1e9
    report("attn fwd (zorro)", t(lambda: K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, **kw)), gflop=fl)
    do = rnd(Mt, HD); dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, N, device=dev)
    report("attn bwd (zorro)", t(lambda: K.attn_bwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dqkv[:, :HD],
                                                    dqkv[:, HD:2 * HD], dqkv[:, 2 * HD:], delta, **kw)), gflop=2.5 * fl)
if want("slot"):
    q = rnd(Mf, HD); kv = rnd(Mt, 2 * HD); kvm = rnd(Fn, 2 * HD); out = torch.empty(Mf, HD, dtype=bf16, device=dev)
    slotmap = torch.full((3, Fn), -1, dtype=torch.int32, device=dev)
    for m in range(3):
        ix = torch.randperm(Fn, device=dev)[:98].sort().values
        slotmap[m, ix] = torch.arange(98, dtype=torch.int32, device=dev)
    seg = torch.tensor([0, 98, 196, 294, 490], dtype=torch.int32, device=dev)
    kw = dict(B=B, F=Fn, H=H, S=4, n_head=nenc, scale=0.125)
    report("slot attn fwd", t(lambda: K.slot_attn_fwd(q, kv, kvm, slotmap, seg, out, None, **kw)), (Mf * HD * 2 * 2 + Mt * 2 * HD * 2) / 1e9)
    dout = rnd(Mf, HD); dq = torch.empty_like(q); dkv = torch.empty_like(kv); dme = torch.zeros(Fn, 2 * HD, device=dev)
    report("slot attn bwd", t(lambda: K.slot_attn_bwd(q, kv, kvm, slotmap, seg, dout, dq, dkv, dme, **kw)), (Mf * HD * 2 * 3 + Mt * 2 * HD * 2 * 2) / 1e9)
if want("geglu"):
    I = 2048
    u = rnd(Mt, 2 * I); dg = rnd(Mt, I); du = torch.empty_like(u)
    report("geglu bwd", t(lambda: K.geglu_bwd(u, dg, du)), Mt * I * 2 * 5 / 1e9)
if want("cast"):
    x = rnd(Mt, D, dt=f32)
    report("cast f32->bf16", t(lambda: K.cast_bf16(x)), Mt * D * 6 / 1e9)
