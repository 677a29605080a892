import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import oracle
from oracle import OracleConfig
from collections import OrderedDict
from _util import build_model, default_sd, make_inputs, rel, pretrain_loss_ours
cfg = OracleConfig(variant='crossattn', dim=192, depth=3, heads=3, image_size=96, patch=16, dec_dim=64, dec_depth=2, dec_heads=2)
sd = default_sd(cfg)
model = build_model(cfg, sd)
x = make_inputs(cfg, 4, 11, 'cuda')
torch.manual_seed(5)
out = model(x, num_encoded_tokens=50, sample_tasks_uniformly=True)
print('counts', [int((m[0]==0).sum()) for m in out[1].values()])
snap = {t: v.detach().clone() for t, v in out[0].items()}
sd_o = OrderedDict((k, v.cuda().requires_grad_(not (k.endswith(".beta") or k.endswith("pos_emb")))) for k, v in sd.items())
torch.manual_seed(5)
ref = oracle.multimae_forward(sd_o, cfg, x, num_encoded_tokens=50, sample_tasks_uniformly=True)
for t in snap: print('before bwd', t, rel(snap[t], ref[0][t]))
loss = pretrain_loss_ours(out, x, cfg.patch)
torch.cuda.synchronize()
for t in snap: print('after loss fwd', t, rel(out[0][t], snap[t]))
loss.backward()
torch.cuda.synchronize()
for t in snap: print('after bwd', t, rel(out[0][t], snap[t]))
