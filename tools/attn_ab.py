#!/usr/bin/env python
"""Same-box A/B of the tcgen05 attention kernel variants (MMF_ATTN_FWD / MMF_ATTN_DQ / MMF_ATTN_DKV, read per launch by
csrc/attention_tc.cu) at the cfg-2 shape (B=256, H=8, nenc=294 + 196 fusion tokens) for several modality splits.
Each variant is checked against variant 0 of its kernel (the round-1 kernel, itself checked against torch in
tests/test_kernels_gpu.py) and timed with CUDA events.  Usage: python tools/attn_ab.py [fwd 0,1,2] [dq 0,1] [dkv 0,1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from incomplete_multimodal_fusion_b200 import _lib, kernels as K  # noqa: E402

if os.environ.get("MMF_LIB"):   # another build of the library (e.g. scratch/prev_libmmf.so) for same-box A/B runs
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])

bf16 = torch.bfloat16
B, nenc, Fn, H = int(os.environ.get("AB_BATCH", "256")), 294, 196, 8
N = nenc + Fn
Mt = B * N
HD = H * 64
dev = "cuda"
args = sys.argv[1:]
sel = {"fwd": [0, 1, 2, 3, 4, 5, 6], "dq": [0, 1, 2, 3], "dkv": [0], "pf": [0]}
for i in range(0, len(args) - 1, 2):
    sel[args[i]] = [int(v) for v in args[i + 1].split(",")]


def timed(fn, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


torch.manual_seed(0)
qkv = torch.randn(Mt, 3 * HD, device=dev).to(bf16)
do = torch.randn(Mt, HD, device=dev).to(bf16)
SPLITS = ((98, 98, 98), (150, 0, 144), (37, 196, 61), (0, 196, 98))[:int(os.environ.get("AB_SPLITS", "4"))]
for counts in SPLITS:
    bounds = [0, counts[0], counts[0] + counts[1], nenc, N]
    seg = torch.tensor(bounds, dtype=torch.int32, device=dev)
    pairs = sum(c * c for c in counts) + Fn * N
    gf = 4 * 64 * pairs * B * H / 1e9
    kw = dict(B=B, H=H, Nq=N, Nk=N, dh=64, scale=0.125, n_head_q=nenc, n_head_k=nenc, seg=seg, nseg=4)
    print("== split %s: %.1f GFLOP forward over allowed pairs" % (counts, gf))
    ref_o = ref_lse = None
    for v, pf in [(v, pf) for v in sel["fwd"] for pf in sel["pf"]]:
        os.environ["MMF_ATTN_FWD"], os.environ["MMF_ATTN_PF"] = str(v), str(pf)
        o = torch.full((Mt, HD), float("nan"), dtype=bf16, device=dev)
        lse = torch.full((B, H, N), float("nan"), device=dev)
        f = lambda: K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, **kw)
        try:
            ms = timed(f)
        except RuntimeError as e:
            print("  fwd v%d FAILED: %s" % (v, str(e)[:200]))
            raise
        if ref_o is None:
            ref_o, ref_lse = o.clone(), lse.clone()
        print("  fwd v%d pf%-2d %7.3f ms  %6.1f TFLOP/s   rel(o) %.2e  rel(lse) %.2e  finite %s" % (
            v, pf, ms, gf / ms, rel(o, ref_o), rel(lse, ref_lse), bool(torch.isfinite(o.float()).all())))
    os.environ["MMF_ATTN_FWD"], os.environ["MMF_ATTN_PF"] = "0", "0"
    o = torch.empty(Mt, HD, dtype=bf16, device=dev)
    lse = torch.empty(B, H, N, device=dev)
    K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, **kw)
    ref_d = None
    for vq in sel["dq"]:
        for vk, pf in [(vk, pf) for vk in sel["dkv"] for pf in sel["pf"]]:
            os.environ["MMF_ATTN_DQ"], os.environ["MMF_ATTN_DKV"], os.environ["MMF_ATTN_PF"] = str(vq), str(vk), str(pf)
            dqkv = torch.full_like(qkv, float("nan"))
            delta = torch.empty(B, H, N, device=dev)
            f = lambda: K.attn_bwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dqkv[:, :HD], dqkv[:, HD:2 * HD],
                                   dqkv[:, 2 * HD:], delta, **kw)
            ms = timed(f)
            if ref_d is None:
                ref_d = dqkv.clone()
            print("  bwd dq v%d dkv v%d pf%-2d %7.3f ms  %6.1f TFLOP/s   rel(dq) %.2e rel(dk) %.2e rel(dv) %.2e  finite %s" % (
                vq, vk, pf, ms, 2.5 * gf / ms, rel(dqkv[:, :HD], ref_d[:, :HD]), rel(dqkv[:, HD:2 * HD], ref_d[:, HD:2 * HD]),
                rel(dqkv[:, 2 * HD:], ref_d[:, 2 * HD:]), bool(torch.isfinite(dqkv.float()).all())))
