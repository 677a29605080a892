import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
bf16 = torch.bfloat16
M = 125440
def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
a = (torch.randn(M, 768, device='cuda') * 0.1).to(bf16); w = (torch.randn(1536, 768, device='cuda') * 0.1).to(bf16)
out = torch.empty(M, 1536, dtype=bf16, device='cuda')
ms = t(lambda: K.gemm(a, w, out)); print('dbg', os.environ.get('MMF_GEMM_DEBUG', '0'), 'qkv bn256', round(ms, 3), 'ms', round(2 * M * 1536 * 768 / ms / 1e9), 'TF')
ms = t(lambda: K.gemm(a, w, out, block_n=128)); print('dbg', os.environ.get('MMF_GEMM_DEBUG', '0'), 'qkv bn128', round(ms, 3), 'ms', round(2 * M * 1536 * 768 / ms / 1e9), 'TF')
