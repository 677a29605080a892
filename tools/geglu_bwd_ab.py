#!/usr/bin/env python
"""Same-box A/B of the fused GEGLU-backward dgrad GEMM's epilogue options (MMF_GEGLU_BWD_ABL, read per launch) at the cfg-2
FFN shape: du = [dg * gelu(gate) | dg * value * gelu'(gate)], dg = dY . W2 kept in TMEM (act=3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import _lib, kernels as K
if os.environ.get("MMF_LIB"):   # another build of the library (e.g. scratch/prev_libmmf.so) for same-box A/B runs
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])
M, D, I = int(os.environ.get("AB_M", "125440")), 768, 2048
bf16 = torch.bfloat16
dY = (torch.randn(M, D, device="cuda") * 0.5).to(bf16); W2 = (torch.randn(D, I, device="cuda") * 0.1).to(bf16)
u = torch.randn(M, 2 * I, device="cuda").to(bf16)
du = torch.empty(M, 2 * I, dtype=bf16, device="cuda")
ref = None
for v in sys.argv[1:] or ("0", "1", "0", "1"):
    os.environ["MMF_GEGLU_BWD_ABL"] = v
    du.fill_(float("nan"))
    for _ in range(3): K.gemm(dY, W2, du, b_mn=True, act=3, out2=u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): K.gemm(dY, W2, du, b_mn=True, act=3, out2=u)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None: ref = du.clone()
    nbytes = M * (2 * I * 2 * 2 + D * 2)
    print("pf=%s  %.4f ms  %.0f TFLOP/s  %.0f GB/s algorithmic  identical %s" % (v, ms, 2 * M * I * D / ms / 1e9, nbytes / ms / 1e6, torch.equal(du, ref)))
