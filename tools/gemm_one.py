#!/usr/bin/env python
"""One GEMM shape, timed (profiling experiments): python tools/gemm_one.py N K [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
N, Kd = int(sys.argv[1]), int(sys.argv[2])
M = int(sys.argv[3]) if len(sys.argv) > 3 else 125440
a = (torch.randn(M, Kd, device="cuda") * .1).bfloat16(); w = (torch.randn(N, Kd, device="cuda") * .1).bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for _ in range(3): K.gemm(a, w, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): K.gemm(a, w, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"M={M} N={N} K={Kd} 2cta={os.environ.get('MMF_GEMM_2CTA','1')}: {ms:.3f} ms {2*M*N*Kd/ms/1e9:.0f} TFLOP/s")
