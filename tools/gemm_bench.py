#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM on the shapes of the ViT-B cfg-2 step (M = 125,440 token rows)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from incomplete_multimodal_fusion_b200 import kernels as K  # noqa: E402
from incomplete_multimodal_fusion_b200.functions import _wgrad_split  # noqa: E402

bf16, f32 = torch.bfloat16, torch.float32
M = int(sys.argv[1]) if len(sys.argv) > 1 else 125440
dev = "cuda"


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
    return (torch.randn(*s, device=dev) * 0.1).to(dt)


rows = []
for name, N, Kd, kw in [("qkv", 1536, 768, {}), ("outproj+res", 768, 512, {"res": True}), ("ffn1 geglu", 2048, 768, {"geglu": True}),
                        ("ffn2+res", 768, 2048, {"res": True}), ("plain N=768 K=768 bf16", 768, 768, {})]:
    a = rnd(M, Kd)
    if kw.get("geglu"):
        w = rnd(2 * N, Kd); out = torch.empty(M, N, dtype=bf16, device=dev); u = torch.empty(M, 2 * N, dtype=bf16, device=dev)
        ms = t(lambda: K.gemm(a, w, out, act=2, out2=u)); fl = 2 * M * 2 * N * Kd
    elif kw.get("res"):
        w = rnd(N, Kd); res = rnd(M, N, dt=f32); out = torch.empty(M, N, dtype=f32, device=dev)
        ms = t(lambda: K.gemm(a, w, out, residual=res)); fl = 2 * M * N * Kd
    else:
        w = rnd(N, Kd); out = torch.empty(M, N, dtype=bf16, device=dev)
        ms = t(lambda: K.gemm(a, w, out)); fl = 2 * M * N * Kd
    rows.append(("fwd " + name, ms, fl))
for name, N, Kd in [("dh1", 768, 1536), ("do", 512, 768), ("dg", 2048, 768), ("dh2", 768, 4096)]:
    a = rnd(M, Kd); w = rnd(Kd, N); out = torch.empty(M, N, dtype=bf16, device=dev)
    ms = t(lambda: K.gemm(a, w, out, b_mn=True)); rows.append(("dgrad " + name, ms, 2 * M * N * Kd))
for name, NO, KI in [("dWqkv", 1536, 768), ("dWo", 768, 512), ("dW1", 4096, 768), ("dW2", 768, 2048)]:
    dy = rnd(M, NO); x = rnd(M, KI); out = torch.zeros(NO, KI, dtype=f32, device=dev)
    sk = _wgrad_split(M, NO * KI, NO, KI)
    ms = t(lambda: K.gemm(dy, x, out, a_mn=True, b_mn=True, split_k=sk)); rows.append((f"wgrad {name} split{sk}", ms, 2 * M * NO * KI))
# cuBLAS reference points
a = rnd(M, 768); w = rnd(1536, 768)
ms = t(lambda: a @ w.t()); rows.append(("cuBLAS qkv", ms, 2 * M * 1536 * 768))
a = rnd(M, 2048); w = rnd(768, 2048)
ms = t(lambda: a @ w.t()); rows.append(("cuBLAS ffn2 (no residual)", ms, 2 * M * 768 * 2048))
for name, ms, fl in rows:
    print(f"{name:32s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s")
