#!/usr/bin/env python
"""Split-K factors for the weight-gradient GEMMs dW[NO, KI] = dY[M, NO]^T . X[M, KI] (fp32 atomics into a zeroed buffer):
time per launch for a range of factors around functions._wgrad_split's choice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
from incomplete_multimodal_fusion_b200.functions import _wgrad_split
bf16 = torch.bfloat16
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for M, NO, KI in [(125440, 4096, 768), (125440, 768, 2048), (125440, 1536, 768), (125636, 1024, 768), (125440, 768, 512),
                  (50176, 4096, 768), (50176, 768, 2048), (50176, 768, 512), (50176, 512, 768), (50176, 1024, 256), (50176, 256, 1024),
                  (50176, 768, 256), (50176, 256, 256)]:
    dy = (torch.randn(M, NO, device="cuda") * .1).to(bf16); x = (torch.randn(M, KI, device="cuda") * .1).to(bf16)
    dw = torch.zeros(NO, KI, dtype=torch.float32, device="cuda")
    cur = _wgrad_split(M, NO * KI, NO, KI)
    tiles = ((NO + 255) // 256) * ((KI + 255) // 256)
    res = []
    for sk in sorted({max(1, cur // 4), max(1, cur // 2), max(1, (74 + tiles - 1) // tiles), max(1, 74 // tiles), cur, min(32, cur * 2)}):
        ms = t(lambda: K.gemm(dy, x, dw, a_mn=True, b_mn=True, split_k=sk))
        res.append((sk, ms))
    best = min(res, key=lambda r: r[1])
    print("M=%6d %4dx%-4d tiles %2d current split %2d: " % (M, NO, KI, tiles, cur) + "  ".join("%s%d: %.1f us" % ("*" if sk == cur else "", sk, ms * 1e3) for sk, ms in res) +
          "   best %d (%.1f%% faster)" % (best[0], 100 * (dict(res)[cur] / best[1] - 1)))
