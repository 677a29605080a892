#!/usr/bin/env python
"""The decoders' GEMM shapes (M = 50,176 = 256 x 196 query / context rows, dims 256 / 512 / 768 / 1024), each alone: time,
TFLOP/s, algorithmic GB/s, the HBM floor at MEASURED_PEAKS' copy rate and the cuBLAS time of the bare product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
from incomplete_multimodal_fusion_b200.functions import _wgrad_split
bf16, f32 = torch.bfloat16, torch.float32
M = int(os.environ.get("AB_M", "50176")); dev = "cuda"
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
def rnd(*s, dt=bf16): return (torch.randn(*s, device=dev) * 0.1).to(dt)
rows = []
for name, N, Kd, kind in [("fc1 gelu+bias+pre", 1024, 256, "gelu"), ("fc1 bias only", 1024, 256, "bias"), ("fc1 plain", 1024, 256, "plain"),
                          ("fc2 bias+res f32", 256, 1024, "res"), ("fc2 plain", 256, 1024, "plain"), ("proj_ctx f32 out", 256, 768, "f32"),
                          ("kv 256->512", 512, 256, "plain"), ("q/out 256->256", 256, 256, "plain"), ("256->768", 768, 256, "plain"),
                          ("768->512", 512, 768, "plain"), ("512->768", 768, 512, "plain")]:
    a = rnd(M, Kd); w = rnd(N, Kd); b = rnd(N, dt=f32)
    nb = M * Kd * 2 + N * Kd * 2
    if kind == "gelu":
        out = torch.empty(M, N, dtype=bf16, device=dev); pre = torch.empty_like(out); nb += 2 * M * N * 2
        fn = lambda: K.gemm(a, w, out, bias=b, act=1, out2=pre)
    elif kind == "bias":
        out = torch.empty(M, N, dtype=bf16, device=dev); nb += M * N * 2
        fn = lambda: K.gemm(a, w, out, bias=b)
    elif kind == "res":
        out = torch.empty(M, N, dtype=f32, device=dev); res = rnd(M, N, dt=f32); nb += 2 * M * N * 4
        fn = lambda: K.gemm(a, w, out, bias=b, residual=res)
    elif kind == "f32":
        out = torch.empty(M, N, dtype=f32, device=dev); nb += M * N * 4
        fn = lambda: K.gemm(a, w, out, bias=b)
    else:
        out = torch.empty(M, N, dtype=bf16, device=dev); nb += M * N * 2
        fn = lambda: K.gemm(a, w, out)
    ms = t(fn); cub = t(lambda: a @ w.t())
    rows.append(("fwd " + name, N, Kd, ms, 2 * M * N * Kd, nb, cub))
for name, N, Kd in [("dgrad 1024->256", 256, 1024), ("dgrad 256->1024", 1024, 256), ("dgrad 256->256", 256, 256), ("dgrad 768->256", 256, 768), ("dgrad 256->768", 768, 256)]:
    a = rnd(M, Kd); w = rnd(Kd, N); out = torch.empty(M, N, dtype=bf16, device=dev)
    ms = t(lambda: K.gemm(a, w, out, b_mn=True)); cub = t(lambda: a @ w)
    rows.append((name, N, Kd, ms, 2 * M * N * Kd, M * Kd * 2 + N * Kd * 2 + M * N * 2, cub))
for name, NO, KI in [("wgrad 1024x256", 1024, 256), ("wgrad 256x1024", 256, 1024), ("wgrad 256x256", 256, 256), ("wgrad 768x256", 768, 256)]:
    dy = rnd(M, NO); x = rnd(M, KI); out = torch.zeros(NO, KI, dtype=f32, device=dev)
    sk = _wgrad_split(M, NO * KI, NO, KI)
    ms = t(lambda: K.gemm(dy, x, out, a_mn=True, b_mn=True, split_k=sk)); cub = t(lambda: dy.t() @ x)
    rows.append((f"{name} split{sk}", NO, KI, ms, 2 * M * NO * KI, M * (NO + KI) * 2 + NO * KI * 4, cub))
for name, N, Kd, ms, fl, nb, cub in rows:
    print(f"{name:28s} N={N:5d} K={Kd:5d} {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {nb/ms/1e6:7.0f} GB/s  HBM floor {nb/6550e3:6.1f} us   cuBLAS {cub*1e3:7.1f} us")
