#!/usr/bin/env python
"""The decoders' self-attention (mma.sync kernels of csrc/attention.cu): B = 256, 8 heads of 32, N = 196 tokens, no segments.
Forward and backward time per launch; MMF_LIB selects another build of the library for same-box A/B runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import _lib, kernels as K
if os.environ.get("MMF_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])
bf16 = torch.bfloat16
B, H, dh = 256, 8, 32
HD = H * dh
def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for N in (196, 256):
    torch.manual_seed(0)
    qkv = torch.randn(B * N, 3 * HD, device="cuda").to(bf16); do = torch.randn(B * N, HD, device="cuda").to(bf16)
    o = torch.empty(B * N, HD, dtype=bf16, device="cuda"); lse = torch.empty(B, H, N, device="cuda")
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, N, device="cuda")
    kw = dict(B=B, H=H, Nq=N, Nk=N, dh=dh, scale=dh ** -0.5, n_head_q=N, n_head_k=N, seg=None, nseg=0)
    f = lambda: K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, **kw)
    g = lambda: K.attn_bwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dqkv[:, :HD], dqkv[:, HD:2 * HD], dqkv[:, 2 * HD:], delta, **kw)
    tf, tb = t(f), t(g)
    gf = 4 * N * N * dh * B * H / 1e9
    print("N=%d: fwd %.3f ms (%.0f TFLOP/s)  bwd %.3f ms (%.0f TFLOP/s)  checksum o %.6f dqkv %.6f" % (
        N, tf, gf / tf, tb, 2.5 * gf / tb, float(o.float().abs().mean()), float(dqkv.float().abs().mean())))
