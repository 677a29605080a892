"""Per-phase clock totals of one modality key-tile CTA of the attention dK/dV kernel (debug build with -DMMF_ATTN_CLOCKS
at scratch/dbg_libmmf.so, see tools/attn_clocks.py; in that build the forward's instrumentation is compiled in too, so
only the backward is launched between reset and read)."""
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
do = torch.randn(Mt, HD, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, N, device="cuda")
kw = dict(B=B, H=H, Nq=N, Nk=N, dh=64, scale=0.125, n_head_q=nenc, n_head_k=nenc, seg=seg, nseg=4)
K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, **kw)
f = lambda: K.attn_bwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dqkv[:, :HD], dqkv[:, HD:2 * HD], dqkv[:, 2 * HD:], delta, **kw)
for _ in range(3): f()
torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
raw.mmf_debug_attn_clocks(buf, 1)
f(); torch.cuda.synchronize()
raw.mmf_debug_attn_clocks(buf, 0)
names = ["wait s_full (both halves)", "ld + exp + pack + st + arrive (both halves)", "-", "stage next lse/delta", "loop top", "bar.sync + issue next lse/delta loads", "head epilogue (+wait acc)",
         "-", "mma: loop top / between halves", "mma: wait ds_full", "mma: acc_empty + fence", "mma: issue dV, dK (+commit)", "mma: issue next S^T, dP^T (+waits)"]
tot = sum(buf[i] for i in range(7))
print("mma warp instrumented total", sum(buf[i] for i in range(8, 13)))
its = max(int(buf[15]), 1)
print("one modality key-tile CTA of dK/dV, heads x query blocks = %d iterations; total cycles" % its, tot)
for i, n in enumerate(names):
    print(f"  {n:42s} {buf[i]:10d} cycles  {100*buf[i]/max(tot,1):5.1f}%   per iteration {buf[i]/its:8.0f}")
