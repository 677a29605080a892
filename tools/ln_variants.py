"""debug: LayerNorm backward variants timed at the cfg-2 shape (optionally with scratch/dbg_libmmf.so)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from incomplete_multimodal_fusion_b200 import _lib

if len(sys.argv) > 1 and sys.argv[1] == "dbg":
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scratch", "dbg_libmmf.so")
from incomplete_multimodal_fusion_b200 import kernels as K

bf16, f32 = torch.bfloat16, torch.float32
Mt, D = 125440, 768


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


x = torch.randn(Mt, D, device="cuda")
g1 = torch.ones(D, device="cuda")
g2 = torch.ones(D, device="cuda")
y = torch.empty(Mt, D, dtype=bf16, device="cuda")
st = torch.empty(Mt, 4, device="cuda")
K.layernorm_fwd(x, g1, y, g2=g2, stats=st)
dy = torch.randn(Mt, D, device="cuda").bfloat16()
dres = torch.randn(Mt, D, device="cuda")
dx = torch.empty(Mt, D, device="cuda")
dxb = torch.empty(Mt, D, dtype=bf16, device="cuda")
dg1 = torch.zeros(D, device="cuda")
dg2 = torch.zeros(D, device="cuda")
ms = t(lambda: K.layernorm_bwd(dy, x, g1, st, dx, dg1, g2=g2, dres=dres, dx_bf16=dxb, dg2=dg2))
print("full            %.3f ms %5.0f GB/s" % (ms, Mt * D * 16 / ms / 1e6))
ms = t(lambda: K.layernorm_bwd(dy, x, g1, st, dx, dg1, g2=g2, dx_bf16=dxb, dg2=dg2))
print("no dres         %.3f ms %5.0f GB/s" % (ms, Mt * D * 12 / ms / 1e6))
ms = t(lambda: K.layernorm_bwd(dy, x, g1, st, dx, dg1, g2=g2, dres=dres, dg2=dg2))
print("no bf16 copy    %.3f ms %5.0f GB/s" % (ms, Mt * D * 14 / ms / 1e6))
ms = t(lambda: K.layernorm_bwd(dy, x, g1, st, dx, dg1, dres=dres, dx_bf16=dxb))
print("single LN       %.3f ms %5.0f GB/s" % (ms, Mt * D * 16 / ms / 1e6))
ms = t(lambda: K.layernorm_fwd(x, g1, y, g2=g2, stats=st))
print("fwd ln2         %.3f ms %5.0f GB/s" % (ms, Mt * D * 6 / ms / 1e6))
delta = torch.randn(Mt, D, device="cuda").bfloat16()
xo = torch.empty(Mt, D, device="cuda")
ms = t(lambda: K.layernorm_fwd(x, g1, y, g2=g2, stats=st, delta=delta, xout=xo))
print("fwd ln2 + delta %.3f ms %5.0f GB/s" % (ms, Mt * D * 12 / ms / 1e6))
