"""The pre-training step around the hot path: model assembly (reference pretrain_mmae.py:45-72,188-248),
loss assembly (:476-500) and the optimiser settings (utils/optim_factory.py:138-176: AdamW, betas
(0.9, 0.95), weight decay 0.05 on every parameter, lr = blr * global_batch / 256).

Data parallelism (reference: DistributedDataParallel, pretrain_mmae.py:342-345): one process per GPU,
identical replicas, batch sharded; `GradAllReduce` packs the gradients into contiguous buckets and runs
NCCL all-reduce (mean) on a side stream while the rest of the step proceeds."""
from collections import OrderedDict
from typing import Dict, List, Optional

import os

import torch
import torch.distributed as dist

from . import functions as Fn
from .multimae import multimae as _plain
from .multimae import multimae_crossattn as _cross
from .multimae import multimae_lstm_s2dsm as _lstm
from .multimae.criterion import HardNegtive_loss, MaskedL1Loss, MaskedMSELoss, dino_loss_func
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
    if variant == "lstm_s2dsm" and channels is None:          # pretrain_mmae_s2dsm.py DOMAIN_CONF (:45-65); BASELINE
        channels = OrderedDict([("s2", 4), ("dem", 1)])       # config 1 uses the 4-band S2 optical input
    ch = channels or {k: v["channels"] for k, v in DOMAIN_CONF.items()}
    ia = OrderedDict((d, PatchedInputAdapter(num_channels=c, stride_level=1, patch_size_full=patch_size, image_size=image_size))
                     for d, c in ch.items())
    ia["fusion"] = FusionInputAdapter(num_channels=1, stride_level=1, patch_size_full=patch_size, image_size=image_size)
    oa = OrderedDict((d, SpatialOutputAdapter(num_channels=c, stride_level=1, patch_size_full=patch_size, dim_tokens=decoder_dim,
                                              depth=decoder_depth, num_heads=decoder_heads, use_task_queries=True, task=d,
                                              context_tasks=list(ch), image_size=image_size, use_xattn=True))
                     for d, c in ch.items())
    cls = {"crossattn": _cross.MultiMAE, "lstm_s2dsm": _lstm.MultiMAE}.get(variant, _plain.MultiMAE)
    kw = {}
    if variant == "lstm_s2dsm":
        from .multimae.zorro_utils import TokenTypes
        kw["return_token_types"] = (TokenTypes.S2, TokenTypes.DEM, TokenTypes.FUSION)   # pretrain_mmae_s2dsm.py:232-236
    return cls(input_adapters=ia, output_adapters=oa, dim_tokens=dim, depth=depth or dflt_depth, dim_head=64, heads=heads,
               ff_mult=4, num_fusion_tokens=(image_size // patch_size) ** 2, **kw)


class GradAllReduce:
    """Gradient all-reduce (sum) for identical replicas, overlapped with the backward pass (reference: DDP's bucketed
    reducer, pretrain_mmae.py:342-345).

    * `reduce_now(tensors)` is the in-backward hook: EncoderStackFn.backward calls it once per encoder layer with that
      layer's freshly produced weight gradients (~55 MB fp32 for ViT-B).  They are packed into a persistent flat bucket
      and all-reduced on a side stream while the next layer's backward kernels run; the hook returns views of the
      bucket, which autograd then installs as `.grad`, so there is no unpack copy.
    * `finish()` reduces whatever did not go through the hook (embeddings, pooling head, decoders: a few M parameters)
      in one more bucket and makes the current stream wait for all outstanding reductions.
    The mean over ranks comes from scaling the loss by 1/world before backward (PretrainStep), not from an extra pass
    over the gradients.  Parameters that never receive a gradient (the off-task `task_embeddings`, `return_tokens`
    under the DINO-style loss -- SURVEY.md 2.2) simply never show up: no DDP-style unused-parameter search."""

    def __init__(self, params: List[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.stream = None
        self._pool = []          # persistent flat buffers, one per reduce call of a step (call order is static)
        self._call = 0
        self._reduced = set()    # data_ptr of gradients already reduced this step
        self._keep = []          # source tensors kept alive until the side stream has consumed them
        # MMF_ALLREDUCE=end (default): the buckets are filled during backward as usual but all-reduced back to back on the
        # main stream in finish(), after the last backward kernel; MMF_ALLREDUCE=overlap launches each bucket's all-reduce
        # on a side stream as soon as it is full.  Overlapped NCCL kernels cannot share an SM with the persistent GEMM
        # CTAs (registers), so they start at kernel boundaries and delay whole GEMM clusters: measured 135.2 ms (overlap)
        # against 133.0 ms (end) per step at 8 GPUs, 130.6 against 129.2 ms at 2 (profiles/r02_allreduce_modes.json).
        self.deferred = os.environ.get("MMF_ALLREDUCE", "end") == "end"
        self._pending = []       # deferred mode: flat buffers awaiting their all-reduce

    def enabled(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _flat(self, numel, device):
        if self._call == len(self._pool):
            self._pool.append(torch.empty(numel, dtype=torch.float32, device=device))
        flat = self._pool[self._call]
        if flat.numel() != numel or flat.device != device:
            flat = self._pool[self._call] = torch.empty(numel, dtype=torch.float32, device=device)
        self._call += 1
        return flat

    def reduce_now(self, tensors: List[torch.Tensor]) -> List[torch.Tensor]:
        """pack -> all-reduce (async w.r.t. the current stream) -> views of the bucket, in the order given"""
        if not self.enabled() or not tensors:
            return tensors
        dev = tensors[0].device
        sizes = [t.numel() for t in tensors]
        flat = self._flat(sum(sizes), dev)
        views = [v.view(t.shape) for v, t in zip(flat.split(sizes), tensors)]
        if dev.type != "cuda":                      # gloo / CPU (tests): same bucketing, no streams
            torch._foreach_copy_(views, [t.reshape(v.shape) for t, v in zip(tensors, views)])
            dist.all_reduce(flat, group=self.group)
        elif self.deferred:
            torch._foreach_copy_(views, list(tensors))
            self._pending.append(flat)
        else:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            ready = torch.cuda.Event()
            ready.record()
            self._keep.extend(tensors)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                torch._foreach_copy_(views, list(tensors))
                dist.all_reduce(flat, group=self.group)
        self._reduced.update(v.data_ptr() for v in views)
        return views

    def reduce_inplace(self, flat: torch.Tensor, views: List[torch.Tensor]):
        """all-reduce a contiguous buffer that already holds gradients (the encoder backward's per-layer zero arena;
        `views` are the parameter gradients inside it): no packing copy, the views stay the `.grad`s"""
        if not self.enabled() or flat.numel() == 0:
            return
        if flat.device.type != "cuda":
            dist.all_reduce(flat, group=self.group)
        elif self.deferred:
            self._pending.append(flat)
        else:
            if self.stream is None:
                self.stream = torch.cuda.Stream()
            ready = torch.cuda.Event()
            ready.record()
            self._keep.append(flat)      # alive until finish() has joined the side stream
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                torch.cuda.nvtx.range_push("mmf.allreduce")
                dist.all_reduce(flat, group=self.group)
                torch.cuda.nvtx.range_pop()
        self._reduced.update(v.data_ptr() for v in views)

    def finish(self):
        """call after backward(): reduces the remaining gradients; returns with every reduction visible to the current stream"""
        if self.enabled():
            rest = [p for p in self.params if p.grad is not None and p.grad.data_ptr() not in self._reduced]
            if rest:
                views = self.reduce_now([p.grad for p in rest])
                for p, v in zip(rest, views):
                    p.grad = v
            if self._pending:
                torch.cuda.nvtx.range_push("mmf.allreduce")
                cm = getattr(dist, "_coalescing_manager", None) if os.environ.get("MMF_ALLREDUCE_COALESCE", "1") != "0" else None
                if cm is not None and self._pending[0].is_cuda:
                    with cm(group=self.group, device=self._pending[0].device):   # one ncclGroup: a single fused launch
                        for flat in self._pending:
                            dist.all_reduce(flat, group=self.group)
                else:
                    for flat in self._pending:
                        dist.all_reduce(flat, group=self.group)
                torch.cuda.nvtx.range_pop()
                self._pending.clear()
            if self.stream is not None:
                torch.cuda.current_stream().wait_stream(self.stream)
        self._call = 0
        self._reduced.clear()
        self._keep.clear()

    reduce = finish   # earlier name


class PretrainStep:
    """One optimisation step: forward, masked reconstruction + contrastive losses, backward, gradient all-reduce,
    AdamW.  Mirrors the body of train_one_epoch (pretrain_mmae.py:437-517) without the per-step host syncs."""

    def __init__(self, model, num_encoded_tokens: int, patch_size: int = 16, blr: float = 1e-4, global_batch: int = 256,
                 weight_decay: float = 0.05, sample_tasks_uniformly: bool = True, alphas: float = 1.0,
                 contrastive_weight: float = 0.3, max_grad_norm: Optional[float] = None, torch_optimizer: bool = False,
                 standardize_depth: bool = False):
        self.model = model
        self.standardize_depth = standardize_depth   # pretrain_mmae.py:87-89, 452-459 (off by default there too)
        self.nenc = num_encoded_tokens
        self.uniformly = sample_tasks_uniformly
        self.alphas = alphas
        self.cw = contrastive_weight
        self.losses = {d: DOMAIN_CONF[d]["loss"](patch_size=patch_size, stride=1) for d in DOMAIN_CONF}
        self.hard_negative = HardNegtive_loss()
        if torch_optimizer:   # torch's own fused AdamW (kept for A/B runs and the CPU tests)
            self.opt = torch.optim.AdamW(model.parameters(), lr=blr * global_batch / 256, betas=(0.9, 0.95),
                                         weight_decay=weight_decay, fused=True)
        else:
            from .optim import FusedAdamW
            self.opt = FusedAdamW(model.parameters(), lr=blr * global_batch / 256, betas=(0.9, 0.95), weight_decay=weight_decay,
                                  max_grad_norm=max_grad_norm)
        self.reducer = GradAllReduce(list(model.parameters()))
        # zero arena for the gradient accumulators outside the encoder stack: every parameter that is not an encoder-block
        # tensor, twice over (transient accumulators such as padded weight gradients share it)
        enc = {id(p) for n, p in model.named_parameters() if n.startswith("blocks.") or n.startswith("fus_blocks.")}
        self._arena_elems = 2 * sum(p.numel() + 4 for p in model.parameters() if p.requires_grad and id(p) not in enc) + 1024
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            model.grad_hook = self.reducer.reduce_now   # per-layer reduction from inside the encoder backward
            model.grad_hook_inplace = self.reducer.reduce_inplace

    def loss(self, out, targets):
        preds, masks = out[0], out[1]
        total = 0
        for d, p in preds.items():
            total = total + self.losses[d](p, targets[d], mask=masks.get(d))
        if len(out) == 8:   # crossattn variant: DINO-style terms, fusion pool = teacher (pretrain_mmae.py:489-493)
            pooled = torch.chunk(out[2], 4, dim=1)
            total = total + self.cw * sum(dino_loss_func(out[5 + i].squeeze(1), pooled[i].squeeze(1)) for i in range(3))
        elif getattr(self.model, "LSTM_FUSION", False):   # s2dsm script: hard-negative pairs, weight 1 (pretrain_mmae_s2dsm.py:482-492)
            a, b, c = [t.squeeze(1) for t in torch.chunk(out[2], 3, dim=1)]
            total = total + self.hard_negative(a, b) + self.hard_negative(a, c) + self.hard_negative(b, c)
        return total

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        # NVTX ranges: forward / loss / backward (the per-layer all-reduces are issued from inside it, range
        # "allreduce" on their side stream's launches) / allreduce-finish / optimizer show up as rows in nsys / ncu timelines
        nvtx = torch.cuda.nvtx
        self.opt.zero_grad(set_to_none=True)
        if self.standardize_depth and 'dem' in inputs:
            from .utils.multimodal_dfc2023 import standardize_depth
            inputs = dict(inputs, dem=standardize_depth(inputs['dem']))
        nvtx.range_push("mmf.forward")
        out = self.model(inputs, num_encoded_tokens=self.nenc, alphas=self.alphas, sample_tasks_uniformly=self.uniformly)
        nvtx.range_pop()
        nvtx.range_push("mmf.loss")
        loss = self.loss(out, inputs)
        nvtx.range_pop()
        Fn.begin_step_arena(self._arena_elems, loss.device)
        nvtx.range_push("mmf.backward")
        try:
            (loss / self.world if self.world > 1 else loss).backward()   # 1/world here -> the all-reduce is a plain sum
        finally:
            Fn.end_step_arena()
            nvtx.range_pop()
        nvtx.range_push("mmf.allreduce_finish")
        self.reducer.finish()
        nvtx.range_pop()
        nvtx.range_push("mmf.optimizer")
        self.opt.step()
        nvtx.range_pop()
        return loss.detach()
