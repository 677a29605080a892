"""The optimiser side of the training step, on the library's own kernels.

Reference: utils/optim_factory.py:138-176 (`create_optimizer`: torch.optim.AdamW, betas (0.9, 0.95), weight decay
0.05 on every parameter through the dict branch), utils/native_scaler.py:20-82 (`NativeScalerWithGradNormCount`:
backward, one `torch.norm` per parameter for the gradient norm, optional `clip_grad_norm_`, `optimizer.step()`) and
the per-iteration cosine learning-rate / weight-decay tables of `utils.cosine_scheduler` consumed by
`train_one_epoch` (pretrain_mmae.py:441-445).

`FusedAdamW.step()` is two launches for the whole model: `mmf_grad_norm` (only when a norm or clipping is asked for;
the clip coefficient stays on the device) and `mmf_adamw_step`, which also refreshes the bf16 weight images the next
forward's GEMMs read, so the per-step fp32 -> bf16 casts (`functions.WEIGHTS`) drop out.  There is no CPU path.
"""
import ctypes as C
import math
from typing import Iterable, List, Optional

import numpy as np
import torch

from . import _lib, functions
from .kernels import _L, _stream, check

CHUNK = 65536   # elements per CTA


def cosine_scheduler(base_value: float, final_value: float, epochs: int, niter_per_ep: int, warmup_epochs: int = 0,
                     start_warmup_value: float = 0.0, warmup_steps: int = -1) -> np.ndarray:
    """Per-iteration schedule table: linear warm-up, then half a cosine from base_value to final_value (the table
    `train_one_epoch` indexes with the global iteration, pretrain_mmae.py:441-445; restated from utils.cosine_scheduler)."""
    warmup_iters = warmup_steps if warmup_steps > 0 else warmup_epochs * niter_per_ep
    warm = np.linspace(start_warmup_value, base_value, warmup_iters) if warmup_iters > 0 else np.array([])
    iters = np.arange(epochs * niter_per_ep - warmup_iters)
    rest = np.array([final_value + 0.5 * (base_value - final_value) * (1 + math.cos(math.pi * i / len(iters))) for i in iters])
    table = np.concatenate((warm, rest))
    assert len(table) == epochs * niter_per_ep
    return table


class FusedAdamW:
    """AdamW over a list of fp32 CUDA parameters, one launch per step.

    `param_groups` mirrors torch's (a single group: `lr`, `weight_decay`, `betas`, `eps`), so schedule code that
    assigns `group["lr"] = table[it]` works unchanged.  `max_grad_norm` > 0 clips the global gradient norm like
    `clip_grad_norm_`; `track_grad_norm` keeps the norm of the last step in `self.grad_norm` (a device scalar)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.05, max_grad_norm: Optional[float] = None, track_grad_norm: bool = False,
                 emit_bf16: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FusedAdamW got no trainable parameters")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
        self.device = self.params[0].device
        # `lr_scale` is read by the reference loop's schedule code (pretrain_mmae.py:441-445: lr_table[it] * group["lr_scale"])
        self.param_groups = [dict(params=self.params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, lr_scale=1.0)]
        self.max_grad_norm = max_grad_norm
        self.track_grad_norm = track_grad_norm or bool(max_grad_norm)
        self.emit_bf16 = emit_bf16
        sizes = [p.numel() for p in self.params]
        offs = np.concatenate(([0], np.cumsum([(n + 3) // 4 * 4 for n in sizes])))   # 16-byte aligned slices
        self._m = torch.zeros(int(offs[-1]), dtype=torch.float32, device=self.device)
        self._v = torch.zeros_like(self._m)
        self.exp_avg = [self._m[int(o):int(o) + n].view_as(p) for o, n, p in zip(offs, sizes, self.params)]
        self.exp_avg_sq = [self._v[int(o):int(o) + n].view_as(p) for o, n, p in zip(offs, sizes, self.params)]
        self.steps = [0] * len(self.params)
        # launch block -> (tensor, chunk)
        ct, ci = [], []
        for t, n in enumerate(sizes):
            for c in range((n + CHUNK - 1) // CHUNK):
                ct.append(t)
                ci.append(c)
        self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=self.device)
        self._chunk_index = torch.tensor(ci, dtype=torch.int32, device=self.device)
        self._nchunks = len(ct)
        nbytes = C.sizeof(_lib.AdamWTensor) * len(self.params)
        self._host = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]   # alternate: the async copy of
        self._tables = [(_lib.AdamWTensor * len(self.params)).from_buffer(h.numpy()) for h in self._host]  # step k may still be queued
        self._copied = [None, None]   # event recorded after each pinned table's H2D copy: waited for before the table is rewritten
        self._dev = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self._scal = torch.zeros(3, dtype=torch.float32, device=self.device)   # sqnorm, norm, clip coefficient
        self._it = 0
        for tab in self._tables:
            for e, p, m, v in zip(tab, self.params, self.exp_avg, self.exp_avg_sq):
                e.p, e.m, e.v = p.data_ptr(), m.data_ptr(), v.data_ptr()

    @property
    def grad_norm(self) -> torch.Tensor:
        return self._scal[1]

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _images(self):
        """bf16 weight images (functions.WEIGHTS) that are plain row-major casts of one of our parameters:
        param index -> [(image base pointer, row pitch)], plus the cache entries they belong to"""
        index = {id(p): i for i, p in enumerate(self.params)}
        per_param = {}
        entries = []
        for key, (refs, _vers, img) in list(functions.WEIGHTS._store.items()):
            srcs = [r() for r in refs]
            if any(s is None or id(s) not in index for s in srcs) or img.dim() != 2 or img.dtype != torch.bfloat16:
                continue
            kind = key[0]
            if kind == "geglu":
                w = srcs[0]
                if img.shape[0] != w.shape[0]:          # padded value | gate halves: leave to the lazy rebuild
                    continue
            elif kind not in ("w", "cat", "hcat"):
                continue
            row, col, ok, slots = 0, 0, True, []
            for s in srcs:
                cols = s.numel() // s.shape[0]
                if kind == "hcat":                       # column-wise concatenation: same rows, a column offset per source
                    if col + cols > img.shape[1] or s.shape[0] != img.shape[0]:
                        ok = False
                        break
                    slots.append((index[id(s)], img.data_ptr() + col * 2, img.stride(0)))
                    col += cols
                    continue
                if cols > img.shape[1] or row + s.shape[0] > img.shape[0]:
                    ok = False
                    break
                slots.append((index[id(s)], img.data_ptr() + row * img.stride(0) * 2, img.stride(0)))
                row += s.shape[0]
            if not ok or any(len(per_param.get(i, [])) >= 2 for i, _, _ in slots):
                continue
            for i, ptr, pitch in slots:
                per_param.setdefault(i, []).append((ptr, pitch))
            entries.append((key, srcs))
        return per_param, entries

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        lr, (b1, b2), eps, wd = float(g["lr"]), g["betas"], float(g["eps"]), float(g["weight_decay"])
        images, entries = self._images() if self.emit_bf16 else ({}, [])
        slot = self._it & 1
        tab = self._tables[slot]
        host = self._host[slot]
        self._it += 1
        if self._copied[slot] is not None:      # the copy queued two steps ago must have read this table before it is rewritten
            self._copied[slot].synchronize()
        active = []
        for i, (e, p) in enumerate(zip(tab, self.params)):
            gr = p.grad
            if gr is None:
                e.n = 0
                continue
            if gr.dtype != torch.float32 or not gr.is_contiguous() or gr.numel() != p.numel() or not gr.is_cuda:
                gr = p.grad = gr.float().contiguous()
            self.steps[i] += 1
            e.g, e.n = gr.data_ptr(), p.numel()
            e.bias_correction1 = 1.0 - b1 ** self.steps[i]
            e.bias_correction2 = 1.0 - b2 ** self.steps[i]
            im = images.get(i, [])
            e.cols = p.numel() // p.shape[0] if (im and p.dim() >= 1) else 0
            e.w16a, e.pitch16a = im[0] if len(im) > 0 else (None, 0)
            e.w16b, e.pitch16b = im[1] if len(im) > 1 else (None, 0)
            active.append(p)
        if not active:
            return
        self._dev.copy_(host, non_blocking=True)
        if self._copied[slot] is None:
            self._copied[slot] = torch.cuda.Event()
        self._copied[slot].record()
        st = _stream()
        L = _L()
        scale_ptr = None
        if self.track_grad_norm:
            check(L.mmf_grad_norm(self._dev.data_ptr(), self._chunk_tensor.data_ptr(), self._chunk_index.data_ptr(), self._nchunks,
                                  CHUNK, float(self.max_grad_norm or 0.0), self._scal[0:].data_ptr(), self._scal[1:].data_ptr(),
                                  self._scal[2:].data_ptr(), st), "mmf_grad_norm")
            if self.max_grad_norm:
                scale_ptr = self._scal[2:].data_ptr()
        check(L.mmf_adamw_step(self._dev.data_ptr(), self._chunk_tensor.data_ptr(), self._chunk_index.data_ptr(), self._nchunks,
                               CHUNK, lr, b1, b2, eps, wd, scale_ptr, st), "mmf_adamw_step")
        # the kernel wrote through raw pointers: tell autograd / the weight cache that the parameters changed ...
        for p in active:
            torch.autograd.graph.increment_version(p)
        # ... and that the images refreshed in the same launch are current
        # (only images whose sources were ALL updated by this launch: an image of a parameter without a gradient was not
        # rewritten and keeps its old stamp)
        act = {id(p) for p in active}
        for key, srcs in entries:
            if not all(id(s) in act for s in srcs):
                continue
            refs, _vers, img = functions.WEIGHTS._store[key]
            functions.WEIGHTS._store[key] = (refs, tuple((s.data_ptr(), s._version) for s in srcs), img)

    # ---- checkpointing: torch.optim.AdamW's layout (the reference saves / restores optimizer.state_dict(),
    # utils/checkpoint.py:83,128) ----
    def state_dict(self):
        state = {}
        for i, p in enumerate(self.params):
            if self.steps[i] > 0:
                state[i] = {"step": torch.tensor(float(self.steps[i])), "exp_avg": self.exp_avg[i].detach().clone(),
                            "exp_avg_sq": self.exp_avg_sq[i].detach().clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        # the keys torch.optim.AdamW expects in a loaded group (its step() indexes them without defaults)
        for k, v in (("amsgrad", False), ("maximize", False), ("foreach", None), ("capturable", False), ("differentiable", False),
                     ("fused", None), ("decoupled_weight_decay", True)):
            groups[0].setdefault(k, v)
        groups[0]["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.params):
            raise ValueError("loaded state dict has %d parameters, the optimiser %d" % (len(ids), len(self.params)))
        pos = {pid: j for j, pid in enumerate(ids)}           # checkpoint parameter id -> our index (torch's order)
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = tuple(v) if k == "betas" else v
        self._m.zero_()
        self._v.zero_()
        self.steps = [0] * len(self.params)
        for pid, st in sd["state"].items():
            j = pos[int(pid)]
            if st["exp_avg"].shape != self.params[j].shape:
                raise ValueError("state of parameter %d has shape %s, expected %s" % (j, tuple(st["exp_avg"].shape), tuple(self.params[j].shape)))
            self.exp_avg[j].copy_(st["exp_avg"])
            self.exp_avg_sq[j].copy_(st["exp_avg_sq"])
            self.steps[j] = int(float(st["step"]))
