"""Per-phase clock totals of one fusion-tile CTA of attn_fwd_tc2_kernel: the softmax warp (slots 0-6) and the MMA-issuing warp
(slots 8-12).  Needs the debug build scratch/dbg_libmmf.so (-DMMF_ATTN_CLOCKS, see tools/attn_clocks.py).
Usage: MMF_ATTN_FWD=<variant> python tools/attn_clocks2.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scratch", "dbg_libmmf.so")
from incomplete_multimodal_fusion_b200 import kernels as K
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
B, nenc, Fn, H = 256, 294, 196, 8
N = nenc + Fn; Mt = B * N; HD = 512
seg = torch.tensor([0, 98, 196, 294, 490], dtype=torch.int32, device="cuda")
qkv = torch.randn(Mt, 3 * HD, device="cuda").bfloat16(); o = torch.empty(Mt, HD, dtype=torch.bfloat16, device="cuda"); lse = torch.empty(B, H, N, device="cuda")
kw = dict(B=B, H=H, Nq=N, Nk=N, dh=64, scale=0.125, n_head_q=nenc, n_head_k=nenc, seg=seg, nseg=4)
f = lambda: K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, **kw)
for variant in sys.argv[1:] or ["1"]:
    os.environ["MMF_ATTN_FWD"] = variant
    for _ in range(3): f()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    raw.mmf_debug_attn_clocks(buf, 1)
    f(); torch.cuda.synchronize()
    raw.mmf_debug_attn_clocks(buf, 0)
    its = max(int(buf[15]), 1)
    names = {0: "softmax: wait s_full", 1: "softmax: LDTM + row max", 2: "softmax: (rescale) exp + pack + STTM issue", 3: "softmax: wait_st + fence + arrive",
             4: "softmax: loop top", 5: "softmax: wait o_full", 6: "softmax: head epilogue", 8: "mma: loop top", 9: "mma: wait p_full A",
             10: "mma: wait p_full B", 11: "mma: o_empty + fence", 12: "mma: issue P.V (+commits)"}
    tot_s = sum(buf[i] for i in range(7)); tot_m = sum(buf[i] for i in range(8, 13))
    print("variant %s: one fusion-tile CTA, %d key blocks; softmax warp %d clk (%.0f / block), mma warp %d clk in the instrumented parts (rest = S issue)" % (variant, its, tot_s, tot_s / its, tot_m))
    for i, n in names.items():
        print(f"  {n:46s} {buf[i]:10d} clk   per block {buf[i]/its:8.0f}")
