#!/usr/bin/env python
"""FFN-1 GEGLU GEMM timed alone next to a plain bf16 GEMM of the same MMA work and cuBLAS (profiling experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import _lib, kernels as K
if os.environ.get("MMF_LIB"):   # e.g. scratch/dbg_libmmf.so built from another revision
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])
M, D, I = 125440, 768, 2048
a = (torch.randn(M, D, device="cuda") * .1).bfloat16(); w = (torch.randn(2 * I, D, device="cuda") * .1).bfloat16()
g = torch.empty(M, I, dtype=torch.bfloat16, device="cuda"); u = torch.empty(M, 2 * I, dtype=torch.bfloat16, device="cuda")
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20
fl = 2 * M * 2 * I * D / 1e9
mode = sys.argv[1] if len(sys.argv) > 1 else "geglu"
if mode == "geglu":
    ms = t(lambda: K.gemm(a, w, g, act=2, out2=u))
elif mode == "geglu_noU":
    ms = t(lambda: K.gemm(a, w, g, act=2))
elif mode == "bwd":      # fused GEGLU backward (dgrad of FFN-2 + elementwise) next to the two-kernel form
    dy = (torch.randn(M, D, device="cuda") * .1).bfloat16(); w2 = (torch.randn(D, I, device="cuda") * .1).bfloat16()
    u.normal_(); du = torch.empty_like(u)
    ms = t(lambda: K.gemm(dy, w2, du, b_mn=True, act=3, out2=u))
    ms2 = t(lambda: (K.gemm(dy, w2, g, b_mn=True), K.geglu_bwd(u, g, du)))
    fl = 2 * M * I * D / 1e9
    print(f"unfused   : {ms2:.3f} ms")
elif mode == "plain":
    ms = t(lambda: K.gemm(a, w, u))
else:
    ms = t(lambda: torch.matmul(a, w.t(), out=u))
print(f"{mode:10s}: {ms:.3f} ms {fl/ms:.0f} TFLOP/s")
