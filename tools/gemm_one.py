import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from incomplete_multimodal_fusion_b200 import kernels as K
bf16 = torch.bfloat16
M = 125440
a = (torch.randn(M, 768, device='cuda') * 0.1).to(bf16); w = (torch.randn(1536, 768, device='cuda') * 0.1).to(bf16)
out = torch.empty(M, 1536, dtype=bf16, device='cuda')
for _ in range(4): K.gemm(a, w, out)
torch.cuda.synchronize(); print('ok')
