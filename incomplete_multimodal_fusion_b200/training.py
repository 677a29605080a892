"""The pre-training step around the hot path: model assembly (reference pretrain_mmae.py:45-72,188-248),
loss assembly (:476-500) and the optimiser settings (utils/optim_factory.py:138-176: AdamW, betas
(0.9, 0.95), weight decay 0.05 on every parameter, lr = blr * global_batch / 256).

Data parallelism (reference: DistributedDataParallel, pretrain_mmae.py:342-345): one process per GPU,
identical replicas, batch sharded; `GradAllReduce` packs the gradients into contiguous buckets and runs
NCCL all-reduce (mean) on a side stream while the rest of the step proceeds."""
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from .multimae import multimae as _plain
from .multimae import multimae_crossattn as _cross
from .multimae.criterion import MaskedL1Loss, MaskedMSELoss, dino_loss_func
from .multimae.input_adapters import FusionInputAdapter, PatchedInputAdapter
from .multimae.output_adapters_simple import SpatialOutputAdapter

# channels / losses per domain: pretrain_mmae.py DOMAIN_CONF (:45-72)
DOMAIN_CONF = OrderedDict([
    ("s1", dict(channels=1, loss=MaskedMSELoss)),
    ("s2", dict(channels=3, loss=MaskedMSELoss)),
    ("dem", dict(channels=1, loss=MaskedL1Loss)),
])
SIZES = {"tiny": (192, 12, 3), "small": (384, 12, 6), "base": (768, 12, 8), "large": (1024, 24, 8)}


def build_pretrain_model(size: str = "base", variant: str = "crossattn", image_size: int = 224, patch_size: int = 16,
                         decoder_dim: int = 256, decoder_depth: int = 2, decoder_heads: int = 8, depth: Optional[int] = None,
                         channels: Optional[Dict[str, int]] = None):
    """MultiMAE with the three PatchedInputAdapters, the fusion adapter and the simple decoders, wired the way
    pretrain_mmae.py:get_model does (but honouring `size`; the reference hard-codes the tiny factory)."""
    dim, dflt_depth, heads = SIZES[size]
    ch = channels or {k: v["channels"] for k, v in DOMAIN_CONF.items()}
    ia = OrderedDict((d, PatchedInputAdapter(num_channels=c, stride_level=1, patch_size_full=patch_size, image_size=image_size))
                     for d, c in ch.items())
    ia["fusion"] = FusionInputAdapter(num_channels=1, stride_level=1, patch_size_full=patch_size, image_size=image_size)
    oa = OrderedDict((d, SpatialOutputAdapter(num_channels=c, stride_level=1, patch_size_full=patch_size, dim_tokens=decoder_dim,
                                              depth=decoder_depth, num_heads=decoder_heads, use_task_queries=True, task=d,
                                              context_tasks=list(ch), image_size=image_size, use_xattn=True))
                     for d, c in ch.items())
    cls = _cross.MultiMAE if variant == "crossattn" else _plain.MultiMAE
    return cls(input_adapters=ia, output_adapters=oa, dim_tokens=dim, depth=depth or dflt_depth, dim_head=64, heads=heads,
               ff_mult=4, num_fusion_tokens=(image_size // patch_size) ** 2)


class GradAllReduce:
    """Bucketed gradient all-reduce (mean) for identical replicas.  Parameters that never receive a gradient
    (the off-task `task_embeddings`, `return_tokens` under the DINO-style loss -- SURVEY.md 2.2) are excluded
    statically after the first step instead of DDP's per-step unused-parameter search."""

    def __init__(self, params: List[torch.nn.Parameter], bucket_mb: int = 64, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.bucket_bytes = bucket_mb << 20
        self.group = group
        self.buckets = None
        self.stream = None

    def _build(self):
        live = [p for p in self.params if p.grad is not None]
        self.buckets = []
        cur, size = [], 0
        for p in reversed(live):                      # reverse registration order ~ order gradients become ready
            cur.append(p)
            size += p.numel() * 4
            if size >= self.bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.empty(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device) for b in self.buckets]

    def reduce(self):
        """call after backward(); returns when the averaged gradients are visible to the current stream"""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if self.buckets is None:
            self._build()
        world = dist.get_world_size(self.group)
        on_gpu = self.flat[0].is_cuda if self.flat else False

        def run():
            for bucket, flat in zip(self.buckets, self.flat):
                sizes = [p.numel() for p in bucket]
                torch._foreach_copy_(list(flat.split(sizes)), [p.grad.reshape(-1) for p in bucket])
                flat.div_(world)
                dist.all_reduce(flat, group=self.group)
                torch._foreach_copy_([p.grad.reshape(-1) for p in bucket], list(flat.split(sizes)))

        if not on_gpu:      # gloo / CPU (tests): same bucketing, no streams
            run()
            return
        if self.stream is None:
            self.stream = torch.cuda.Stream()
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            run()
        cur.wait_stream(self.stream)


class PretrainStep:
    """One optimisation step: forward, masked reconstruction + contrastive losses, backward, gradient all-reduce,
    AdamW.  Mirrors the body of train_one_epoch (pretrain_mmae.py:437-517) without the per-step host syncs."""

    def __init__(self, model, num_encoded_tokens: int, patch_size: int = 16, blr: float = 1e-4, global_batch: int = 256,
                 weight_decay: float = 0.05, sample_tasks_uniformly: bool = True, alphas: float = 1.0,
                 contrastive_weight: float = 0.3):
        self.model = model
        self.nenc = num_encoded_tokens
        self.uniformly = sample_tasks_uniformly
        self.alphas = alphas
        self.cw = contrastive_weight
        self.losses = {d: DOMAIN_CONF[d]["loss"](patch_size=patch_size, stride=1) for d in DOMAIN_CONF}
        self.opt = torch.optim.AdamW(model.parameters(), lr=blr * global_batch / 256, betas=(0.9, 0.95),
                                     weight_decay=weight_decay, fused=True)
        self.reducer = GradAllReduce(list(model.parameters()))

    def loss(self, out, targets):
        preds, masks = out[0], out[1]
        total = 0
        for d, p in preds.items():
            total = total + self.losses[d](p, targets[d], mask=masks.get(d))
        if len(out) == 8:   # crossattn variant: DINO-style terms, fusion pool = teacher (pretrain_mmae.py:489-493)
            pooled = torch.chunk(out[2], 4, dim=1)
            total = total + self.cw * sum(dino_loss_func(out[5 + i].squeeze(1), pooled[i].squeeze(1)) for i in range(3))
        return total

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        out = self.model(inputs, num_encoded_tokens=self.nenc, alphas=self.alphas, sample_tasks_uniformly=self.uniformly)
        loss = self.loss(out, inputs)
        loss.backward()
        self.reducer.reduce()
        self.opt.step()
        return loss.detach()
