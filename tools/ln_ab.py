#!/usr/bin/env python
"""Same-box A/B of the LayerNorm kernels' L2 prefetch distance (MMF_LN_BWD_L2PF / MMF_LN_FWD_L2PF, row steps ahead,
0 = off; read per launch) at the cfg-2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
Mt, D = 125440, 768
x = torch.randn(Mt, D, device="cuda"); g1 = torch.rand(D, device="cuda") + 0.5; g2 = torch.rand(D, device="cuda") + 0.5
y = torch.empty(Mt, D, dtype=torch.bfloat16, device="cuda"); st = torch.empty(Mt, 4, device="cuda")
delta = torch.randn(Mt, D, device="cuda").bfloat16(); xout = torch.empty(Mt, D, device="cuda")
K.layernorm_fwd(x, g1, y, g2=g2, stats=st)
dy = torch.randn(Mt, D, device="cuda").bfloat16(); dres = torch.randn(Mt, D, device="cuda")
dx = torch.empty(Mt, D, device="cuda"); dxb = torch.empty(Mt, D, dtype=torch.bfloat16, device="cuda")
dg1 = torch.zeros(D, device="cuda"); dg2 = torch.zeros(D, device="cuda")
def timed(fn, n=20):
    for _ in range(3): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out
def bwd():
    K.layernorm_bwd(dy, x, g1, st, dx, dg1, g2=g2, dres=dres, dx_bf16=dxb, dg2=dg2)
    return dx, dxb
def fwd():
    K.layernorm_fwd(x, g1, y, g2=g2, stats=st)
    return y, st
def fwd_res():
    K.layernorm_fwd(x, g1, y, g2=g2, stats=st, delta=delta, xout=xout)
    return y, xout
os.environ["MMF_LN_BWD_RING"] = "0"
dg1.zero_(); dg2.zero_(); bwd(); torch.cuda.synchronize()
ref_bwd = (dx.clone(), dxb.clone(), dg1.clone(), dg2.clone())
for ring in ("1", "0", "1"):
    os.environ["MMF_LN_BWD_RING"] = ring
    dg1.zero_(); dg2.zero_(); dx.fill_(float("nan")); dxb.fill_(float("nan")); bwd(); torch.cuda.synchronize()
    rel = lambda a, b: float((a - b).norm() / b.norm())
    chk = "dx identical %s dxb identical %s rel(dg1) %.1e rel(dg2) %.1e" % (torch.equal(dx, ref_bwd[0]), torch.equal(dxb, ref_bwd[1]), rel(dg1, ref_bwd[2]), rel(dg2, ref_bwd[3]))
    ms, _ = timed(bwd)
    print("bwd ring=%s  %.4f ms %.0f GB/s  %s" % (ring, ms, Mt * D * 16 / ms / 1e6, chk))
os.environ["MMF_LN_BWD_RING"] = "0"
for name, fn, env, nbytes in (("bwd", bwd, "MMF_LN_BWD_L2PF", 16), ("fwd", fwd, "MMF_LN_FWD_L2PF", 6), ("fwd+residual", fwd_res, "MMF_LN_FWD_L2PF", 12)):
    ref = None
    for v in ("0", "1", "2", "3", "4", "0", "2"):
        os.environ[env] = v
        ms, out = timed(fn)
        if ref is None: ref = tuple(t.clone() for t in out)
        same = all(torch.equal(a, b) for a, b in zip(out, ref))
        print("%-13s l2pf=%s  %.4f ms %.0f GB/s  identical %s" % (name, v, ms, Mt * D * nbytes / ms / 1e6, same))
    os.environ.pop(env)
