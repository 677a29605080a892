#!/usr/bin/env python
"""Same-box A/B of the LayerNorm backward's L2 prefetch (MMF_LN_BWD_L2PF, read per launch) at the cfg-2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
Mt, D = 125440, 768
x = torch.randn(Mt, D, device="cuda"); g1 = torch.rand(D, device="cuda") + 0.5; g2 = torch.rand(D, device="cuda") + 0.5
y = torch.empty(Mt, D, dtype=torch.bfloat16, device="cuda"); st = torch.empty(Mt, 4, device="cuda")
K.layernorm_fwd(x, g1, y, g2=g2, stats=st)
dy = torch.randn(Mt, D, device="cuda").bfloat16(); dres = torch.randn(Mt, D, device="cuda")
dx = torch.empty(Mt, D, device="cuda"); dxb = torch.empty(Mt, D, dtype=torch.bfloat16, device="cuda")
dg1 = torch.zeros(D, device="cuda"); dg2 = torch.zeros(D, device="cuda")
def run():
    K.layernorm_bwd(dy, x, g1, st, dx, dg1, g2=g2, dres=dres, dx_bf16=dxb, dg2=dg2)
    return dx, dxb, dg1, dg2
ref = None
for v in ("0", "1", "0", "1"):
    os.environ["MMF_LN_BWD_L2PF"] = v
    for _ in range(3): out = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): out = run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    if ref is None: ref = tuple(t.clone() for t in out)
    same = all(torch.equal(a, b) for a, b in zip(out[:2], ref[:2]))
    print("l2pf=%s  %.4f ms %.0f GB/s  dx identical %s" % (v, ms, Mt * D * 16 / ms / 1e6, same))
