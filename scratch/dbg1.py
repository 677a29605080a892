import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle
from oracle import OracleConfig
from _util import build_model, default_sd, make_inputs, rel
from incomplete_multimodal_fusion_b200 import kernels as K
bf16=torch.bfloat16
cfg = OracleConfig(variant='crossattn', dim=192, depth=3, heads=3, image_size=96, patch=16, dec_dim=64, dec_depth=2, dec_heads=2)
sd = default_sd(cfg)
model = build_model(cfg, sd)
x = make_inputs(cfg, 4, 11, 'cuda')
torch.manual_seed(5)
out = model(x, num_encoded_tokens=50, sample_tasks_uniformly=True)
sdc = {k: v.cuda() for k, v in sd.items()}
for t in ('s1','s2','dem'):
    ref = oracle.simple_output_adapter(sdc, t, out[4].detach(), (96,96), cfg)
    print(t, 'decoder-only rel', rel(out[0][t], ref))
# piecewise for s2
ad = model.output_adapters['s2']
import incomplete_multimodal_fusion_b200.functions as Fn
B,N,D = out[4].shape
p='output_adapters.s2.'
xx = Fn.linear(out[4].detach().reshape(B*N,-1), ad.proj_context.weight, ad.proj_context.bias + ad.task_embeddings['s2'].reshape(-1), out_f32=True).view(B,N,64)
xr = torch.nn.functional.linear(out[4].detach(), sdc[p+'proj_context.weight'], sdc[p+'proj_context.bias']) + sdc[p+'task_embeddings.s2']
print('proj_context', rel(xx, xr))
y = ad.decoder_transformer(xx)
yr = xr
for j in range(2): yr = oracle.vit_block(sdc, p+f'decoder_transformer.{j}.', yr, 2, cfg)
print('blocks', rel(y, yr))
z = Fn.linear(y.reshape(B*N,64), ad.out_proj.weight, ad.out_proj.bias)
zr = torch.nn.functional.linear(yr, sdc[p+'out_proj.weight'], sdc[p+'out_proj.bias']).reshape(B*N,-1)
print('out_proj', rel(z, zr), z.shape)
zz = torch.nn.functional.linear(y.reshape(B*N,64).float(), ad.out_proj.weight, ad.out_proj.bias)
print('out_proj vs torch same input', rel(z, zz))
for r0 in (0, 128):
  for c0 in (0,256,512):
    print(r0,c0, rel(z[r0:r0+128, c0:c0+256], zz[r0:r0+128, c0:c0+256]))
img = Fn.UnpatchifyFn.apply(z, B, 3, 96, 96, 16)
print('unpatchify', rel(img, oracle.functional._unpatchify(z.view(B,N,-1), cfg, 3, 96, 96)))
# raw gemm repro
A = torch.randn(144, 64, device='cuda').to(bf16); W = torch.randn(768, 64, device='cuda').to(bf16); b = torch.randn(768, device='cuda')
o = torch.empty(144, 768, dtype=bf16, device='cuda'); K.gemm(A, W, o, bias=b)
print('raw gemm', rel(o, A.float()@W.float().t()+b))
