#!/usr/bin/env python
"""Per-phase clock totals of one CTA pair of the fused GEGLU-backward dgrad GEMM (gemm2_tcgen05_kernel<EPI_GEGLU_BWD>): one
epilogue warp, the MMA-issuing warp and the TMA producer of cluster 5.  Needs a debug build of gemm.cu with -DMMF_GEMM_CLOCKS
linked into scratch/dbg_libmmf.so (the other objects as built):
  nvcc <_build.NVCC_FLAGS> -DMMF_GEMM_CLOCKS -c incomplete_multimodal_fusion_b200/csrc/gemm.cu -o /tmp/dbgobj/gemm.o
  nvcc -shared -o scratch/dbg_libmmf.so /tmp/dbgobj/gemm.o <the other .o of _build.OBJ> -gencode arch=compute_100a,code=sm_100a
Usage: python tools/gemm_clocks.py [MMF_GEGLU_BWD_ABL values ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scratch", "dbg_libmmf.so")
from incomplete_multimodal_fusion_b200 import kernels as K
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
M, D, I = int(os.environ.get("AB_M", "125440")), 768, 2048
bf16 = torch.bfloat16
dY = (torch.randn(M, D, device="cuda") * 0.5).to(bf16); W2 = (torch.randn(D, I, device="cuda") * 0.1).to(bf16)
u = torch.randn(M, 2 * I, device="cuda").to(bf16); du = torch.empty(M, 2 * I, dtype=bf16, device="cuda")
f = lambda: K.gemm(dY, W2, du, b_mn=True, act=3, out2=u)
if os.environ.get("PLAIN") and not os.environ.get("WGRAD"):   # a plain bf16 GEMM instead: PLAIN=N,K[,geglu] (fwd layout), only the mma / tma rows are meaningful
    sp = os.environ["PLAIN"].split(",")
    N_, K_ = int(sp[0]), int(sp[1]); act = 2 if len(sp) > 2 else 0
    a_ = (torch.randn(M, K_, device="cuda") * .1).to(bf16); w_ = (torch.randn(N_, K_, device="cuda") * .1).to(bf16)
    o_ = torch.empty(M, N_ // 2 if act == 2 else N_, dtype=bf16, device="cuda")
    f = lambda: K.gemm(a_, w_, o_, act=act)
if os.environ.get("WGRAD"):   # WGRAD=NO,KI: dW[NO, KI] = dY[M, NO]^T . X[M, KI] (both operands MN-major, split-K, fp32 atomics)
    from incomplete_multimodal_fusion_b200.functions import _wgrad_split
    NO, KI = [int(t) for t in os.environ["WGRAD"].split(",")]
    dy_ = (torch.randn(M, NO, device="cuda") * .1).to(bf16); x_ = (torch.randn(M, KI, device="cuda") * .1).to(bf16)
    dw_ = torch.zeros(NO, KI, dtype=torch.float32, device="cuda")
    sk = int(os.environ.get("SPLITK", _wgrad_split(M, NO * KI, NO, KI)))
    print("wgrad %dx%d K=%d split_k=%d" % (NO, KI, M, sk))
    f = lambda: K.gemm(dy_, x_, dw_, a_mn=True, b_mn=True, split_k=sk)
    os.environ["PLAIN"] = "1"
for v in sys.argv[1:] or ["0"]:
    os.environ["MMF_GEGLU_BWD_ABL"] = v
    for _ in range(3): f()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 32)()
    raw.mmf_debug_gemm_clocks(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    raw.mmf_debug_gemm_clocks(buf, 0)
    tiles = max(int(buf[31]), 1)
    names = {0: "epi: loop top / coords", 1: "epi: wait tmem_full (accumulator)", 2: "epi: tcgen05.ld (+ release on h=1)", 3: "epi: wait boxes h=0",
             4: "epi: wait boxes h=1", 5: "epi: GEGLU-backward math + fence", 6: "epi: issue stores", 7: "epi: wait store read", 8: "epi: issue next loads",
             16: "mma: issue + loop", 17: "mma: wait tmem_empty", 18: "mma: wait full_bar (operands)", 20: "tma: issue", 21: "tma: wait empty_bar (slot free)"}
    if os.environ.get("PLAIN"):
        names = {1: "epi: wait tmem_full (accumulator)", 2: "epi: tcgen05.ld 2 x 64 columns + release", 3: "epi: wait previous stores read",
                 6: "epi: pack box 0", 7: "epi: pack box 1", 8: "epi: fence.proxy.async", 4: "epi: syncwarp", 5: "epi: issue stores (+ loop)", **{k: v for k, v in names.items() if k >= 16}}
    print("flags %s: %.3f ms, %d tiles on this pair, %.0f clk per tile (clock() domain)" % (v, e0.elapsed_time(e1), tiles, sum(buf[i] for i in range(9)) / tiles))
    for i, n in names.items():
        print(f"  {n:44s} {buf[i]:10d} clk   per tile {buf[i]/tiles:8.0f}")
