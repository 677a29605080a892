"""torch.autograd.Function wrappers around the C-ABI kernels (host-side plumbing only).

Numerics contract = the reference under CUDA autocast(bf16): the
residual stream, LayerNorm, softmax statistics and losses are fp32; every contraction takes bf16
operands with fp32 accumulation.  Weight gradients are produced in fp32 straight from the
accumulators.

The encoder stack (all zorro blocks, and for the crossattn variant all Block_Fusion blocks) is ONE
autograd node, `EncoderStackFn`, with a hand-ordered backward: it keeps the token stream in the
planar layout (all modality tokens of the batch, then all fusion tokens), never re-materialises the
concatenated tensor between blocks, and accumulates the residual-stream gradient in place.
"""
import weakref
from typing import List, Optional, Sequence

import os

import torch

from . import kernels as K

bf16, f32 = torch.bfloat16, torch.float32


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


# ------------------------------------------------------------------------------------------------
# bf16 weight images (what autocast's per-forward weight cast does in the reference), cached on the
# parameter's version counter so they are rebuilt once per optimizer step, not once per use
# ------------------------------------------------------------------------------------------------
class _WeightCache:
    """key -> (weak refs to the source parameters, their versions, bf16 image).  Entries are validated by
    object identity through the weak refs: `id()` values are recycled once a model is garbage collected."""

    def __init__(self):
        self._store = {}

    def get(self, key, params: Sequence[torch.Tensor], build):
        hit = self._store.get(key)
        if hit is not None:
            refs, vers, val = hit
            if len(refs) == len(params) and all(r() is p for r, p in zip(refs, params)) and \
                    vers == tuple((p.data_ptr(), p._version) for p in params):
                return val
        val = build()
        if len(self._store) > 4096:
            self._store = {k: v for k, v in self._store.items() if all(r() is not None for r in v[0])}
        self._store[key] = (tuple(weakref.ref(p) for p in params), tuple((p.data_ptr(), p._version) for p in params), val)
        return val

    def clear(self):
        self._store.clear()


WEIGHTS = _WeightCache()


def invalidate_weight_cache():
    """Drop every cached bf16 weight image.  The cache is stamped on (data_ptr, Tensor._version): writes that bypass the
    version counter -- `p.data.copy_()`, `dist.broadcast(p.data, ..)`, EMA / weight surgery through `.data`, a foreign fused
    optimiser -- must be followed by this call (torch optimisers do it through a step hook, `load_state_dict` through the
    model's post hook, see MultiMAEBase.__init__), or the GEMMs keep reading the old images."""
    WEIGHTS.clear()

# torch's fused optimisers (observed: AdamW(fused=True), torch 2.11) update parameters WITHOUT bumping their version
# counters, so version stamps alone would keep serving the pre-update images.  Any torch optimiser step therefore drops
# the cache; `optim.FusedAdamW` (not a torch Optimizer) refreshes the images itself and re-stamps them instead.
try:
    from torch.optim.optimizer import register_optimizer_step_post_hook
    register_optimizer_step_post_hook(lambda _opt, _args, _kwargs: WEIGHTS.clear())
except ImportError:   # pragma: no cover  (older torch: callers must invalidate by hand)
    pass


def w_bf16(w: torch.Tensor, rows_pad: Optional[int] = None, cols_pad: Optional[int] = None) -> torch.Tensor:
    """bf16 image of a 2-D (or conv 4-D, flattened) fp32 weight, zero padded."""
    w2 = w.detach().reshape(w.shape[0], -1)
    return WEIGHTS.get(("w", id(w), rows_pad, cols_pad), [w], lambda: K.cast_bf16(w2, rows_pad=rows_pad, cols_pad=cols_pad))


def w_cat_bf16(ws: Sequence[torch.Tensor]) -> torch.Tensor:
    """row-concatenation of several [n_i, K] weights as one bf16 matrix (to_q || to_kv)."""
    def build():
        rows = sum(w.shape[0] for w in ws)
        out = torch.empty(rows, ws[0].shape[1], dtype=bf16, device=ws[0].device)
        r = 0
        for w in ws:
            K.cast_bf16(w.detach(), out[r:r + w.shape[0]])
            r += w.shape[0]
        return out
    return WEIGHTS.get(("cat",) + tuple(id(w) for w in ws), list(ws), build)


def w_hcat_bf16(ws: Sequence[torch.Tensor], cols_pad: int) -> torch.Tensor:
    """column-concatenation of several [D, K_i] (conv weights flattened) as one bf16 matrix [D, cols_pad], zero padded:
    the B operand of the one-GEMM patch embedding of all modalities (EmbedFn)"""
    def build():
        out = torch.zeros(ws[0].shape[0], cols_pad, dtype=bf16, device=ws[0].device)
        c = 0
        for w in ws:
            w2 = w.detach().reshape(w.shape[0], -1)
            K.cast_bf16(w2, out[:, c:c + w2.shape[1]])
            c += w2.shape[1]
        return out
    return WEIGHTS.get(("hcat", cols_pad) + tuple(id(w) for w in ws), list(ws), build)


def w_geglu_bf16(w1: torch.Tensor, ipad: int) -> torch.Tensor:
    """[2I, D] GEGLU weight -> bf16 [2*ipad, D]: value rows at [0, I), gate rows at [ipad, ipad+I), zero padding."""
    def build():
        I = w1.shape[0] // 2
        if I == ipad:
            return K.cast_bf16(w1.detach())
        out = torch.zeros(2 * ipad, w1.shape[1], dtype=bf16, device=w1.device)
        K.cast_bf16(w1.detach()[:I], out[:I])
        K.cast_bf16(w1.detach()[I:], out[ipad:ipad + I])
        return out
    return WEIGHTS.get(("geglu", id(w1), ipad), [w1], build)


_GEMM_CLUSTERS = 74   # CTA pairs of the persistent GEMM grid on a 148-SM B200


def _wgrad_split(tokens: int, out_elems: int, n_out: Optional[int] = None, k_in: Optional[int] = None) -> int:
    """split-K factor for a weight-gradient GEMM whose contraction runs over `tokens` rows: ONE wave of work items
    (256 x 256 output tiles x splits <= the 74 CTA pairs) when that fills at least 85 % of the grid, else two waves.
    Measured (tools/wgrad_split_ab.py): against the earlier always-two-waves rule the decoders' shapes at 50,176 tokens run
    14-29 % faster (768x512, 1024x256, 768x256, 256x256), 768x512 at 125,440 tokens 8 %, the large ones within 2 %."""
    if os.environ.get("MMF_WGRAD_TWO_WAVES") == "1":   # the earlier rule, for A/B runs
        return max(1, min(max(1, (148 * 2) // max(1, out_elems // (128 * 256))), tokens // 1024 if tokens >= 2048 else 1, 32))
    if n_out and k_in:
        tiles = ((n_out + 255) // 256) * ((k_in + 255) // 256)
    else:
        tiles = max(1, out_elems // (256 * 256))
    s = _GEMM_CLUSTERS // tiles
    if s < 1 or tiles * s < 0.85 * _GEMM_CLUSTERS:
        s = max(1, (2 * _GEMM_CLUSTERS) // tiles)
    return max(1, min(s, tokens // 512 if tokens >= 2048 else 1, _GEMM_CLUSTERS))


# ------------------------------------------------------------------------------------------------
# Zero-initialised fp32 gradient accumulators outside the encoder stack (decoder / pooling / embedding weights, biases,
# LayerNorm gammas: ~100 small tensors per step).  A training step may open an arena (one zero-filled buffer, one
# memset); without one every accumulator is its own torch.zeros as before.
# ------------------------------------------------------------------------------------------------
_STEP_ARENA = None   # (buffer, offset) of the step in flight; module-global: autograd runs backward on its own thread


def begin_step_arena(numel: int, device):
    global _STEP_ARENA
    _STEP_ARENA = [torch.zeros(int(numel), dtype=f32, device=device), 0]


def end_step_arena():
    global _STEP_ARENA
    _STEP_ARENA = None


def zeros_f32(*shape, device):
    a = _STEP_ARENA
    n = 1
    for d in shape:
        n *= int(d)
    if a is None or a[0].device != device or a[1] + n > a[0].numel():
        return torch.zeros(*shape, dtype=f32, device=device)
    v = a[0][a[1]:a[1] + n].view(*shape)
    a[1] += (n + 3) // 4 * 4          # 16-byte aligned views
    return v


def wgrad(dy: torch.Tensor, x: torch.Tensor, out_rows: Optional[int] = None, out_cols: Optional[int] = None,
          out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """dW[n_out, k_in] = dy^T x with dy [T, n_out], x [T, k_in] both bf16 token-major (MN-major operands)."""
    n_out = out_rows or dy.shape[1]
    k_in = out_cols or x.shape[1]
    if out is None:
        out = zeros_f32(n_out, k_in, device=dy.device)
    K.gemm(dy, x, out, a_mn=True, b_mn=True, split_k=_wgrad_split(dy.shape[0], n_out * k_in, n_out, k_in), accumulate=accumulate,
           M=n_out, N=k_in, K=dy.shape[0])
    return out


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    if x.dtype == bf16:
        return x
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    return K.cast_bf16(x2).view(x.shape)


# ------------------------------------------------------------------------------------------------
# generic pieces used by the pooling head and the decoders
# ------------------------------------------------------------------------------------------------
def _split_bf16(x: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi + lo = x to ~16 mantissa bits: lets a bf16 tensor-core GEMM consume an fp32 operand
    without the single 2^-9 rounding (used where one rounded row stands for a whole batch, see LinearFn.precise_grad)"""
    hi = K.cast_bf16(x)
    lo = K.cast_bf16((x - hi.float()).contiguous())
    return hi, lo


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) [+ residual].  x: [M, K] bf16 or f32 (cast once to bf16); y bf16, or f32 when a
    residual (f32) is given or out_f32.  act: 0 none, 1 exact GELU.

    precise_grad: the incoming gradient (fp32) enters the two backward GEMMs as a hi + lo pair of bf16 matrices instead of
    one bf16 rounding.  For batch-INVARIANT rows (the pooling head's learned queries): the reference projects B copies of
    them and rounds B per-sample gradients to bf16 before summing, so its rounding error averages out over the batch; here
    the batch is summed first (fp32) and a single bf16 rounding of that sum would not average (measured 1.3-1.4x the
    reference-autocast error on attn_pool.to_q / return_tokens gradients without this)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, act, out_f32, precise_grad=False):
        xb = to_bf16(x)
        wb = w_bf16(weight)
        M, N = xb.shape[0], weight.shape[0]
        f32_out = out_f32 or residual is not None
        out = torch.empty(M, N, dtype=f32 if f32_out else bf16, device=x.device)
        pre = None
        if act == 1:
            if f32_out:
                raise RuntimeError("GELU epilogue is only provided for bf16 outputs")
            pre = torch.empty(M, N, dtype=bf16, device=x.device)
        if M > 0:
            K.gemm(xb, wb, out, bias=None if bias is None else bias.detach(), act=act,
                   residual=None if residual is None else residual.detach(), out2=pre)
        ctx.save_for_backward(xb, weight, pre)
        ctx.x_dtype = x.dtype
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.precise = bool(precise_grad)
        return out

    @staticmethod
    def backward(ctx, dy):
        xb, weight, pre = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.precise and dy.dtype == f32 and pre is None and not ctx.has_bias and xb.shape[0] > 0:
            hi, lo = _split_bf16(dy)
            wb = w_bf16(weight)
            dx = dw = None
            if ctx.needs_input_grad[0]:
                acc = torch.empty(xb.shape, dtype=f32, device=dy.device)
                K.gemm(hi, wb, acc, b_mn=True)
                K.gemm(lo, wb, acc, b_mn=True, accumulate=True)
                dx = acc if ctx.x_dtype == f32 else acc.to(ctx.x_dtype)
            if ctx.needs_input_grad[1]:
                dw = wgrad(hi, xb)
                wgrad(lo, xb, out=dw, accumulate=True)
                dw = dw.view(weight.shape)
            return dx, dw, None, (dy if (ctx.has_res and ctx.needs_input_grad[3]) else None), None, None, None
        dyb = to_bf16(dy)
        if pre is not None:
            dyb = K.gelu_bwd(pre, dyb, torch.empty_like(dyb))
        wb = w_bf16(weight)
        M, Kd = xb.shape
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, Kd, dtype=ctx.x_dtype, device=dy.device)
            if M > 0:
                K.gemm(dyb, wb, dx, b_mn=True)
        if ctx.needs_input_grad[1]:
            dw = wgrad(dyb, xb) if M > 0 else torch.zeros_like(weight)
            dw = dw.view(weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = zeros_f32(weight.shape[0], device=dy.device)
            if M > 0:
                K.colsum(dyb, db)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return dx, dw, db, dres, None, None, None


def linear(x, weight, bias=None, residual=None, act=0, out_f32=False, precise_grad=False):
    return LinearFn.apply(x, weight, bias, residual, act, out_f32, precise_grad)


class LayerNormFn(torch.autograd.Function):
    """single LayerNorm over the last dim of a 2-D fp32 tensor (gamma, optional bias), bf16 or f32 output"""

    @staticmethod
    def forward(ctx, x, gamma, bias, eps, out_bf16):
        x = x.contiguous()
        rows, D = x.shape
        y = torch.empty(rows, D, dtype=bf16 if out_bf16 else f32, device=x.device)
        stats = torch.empty(rows, 4, dtype=f32, device=x.device)
        K.layernorm_fwd(x, gamma.detach(), y, b1=None if bias is None else bias.detach(), eps1=eps, stats=stats)
        ctx.save_for_backward(x, gamma, bias, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, bias, stats = ctx.saved_tensors
        dy = dy.contiguous()
        D = x.shape[1]
        dx = torch.empty_like(x)
        dg = zeros_f32(D, device=x.device)
        db = zeros_f32(D, device=x.device) if bias is not None else None
        K.layernorm_bwd(dy, x, gamma.detach(), stats, dx, dg, b1=None if bias is None else bias.detach(), db1=db)
        return dx, dg, db, None, None


def layer_norm(x, gamma, bias=None, eps=1e-5, out_bf16=False):
    return LayerNormFn.apply(x, gamma, bias, eps, out_bf16)


class AddLayerNormFn(torch.autograd.Function):
    """(x_new, y) = (x + delta, LN(x + delta)): the residual add of a sub-layer fused into the next LayerNorm
    (x fp32 stream, delta = bf16 output of the sub-layer's last Linear, y bf16)"""

    @staticmethod
    def forward(ctx, x, delta, gamma, bias, eps):
        x = x.contiguous()
        delta = delta.contiguous()
        rows, D = x.shape
        xout = torch.empty(rows, D, dtype=f32, device=x.device)
        y = torch.empty(rows, D, dtype=bf16, device=x.device)
        stats = torch.empty(rows, 4, dtype=f32, device=x.device)
        K.layernorm_fwd(x, gamma.detach(), y, b1=None if bias is None else bias.detach(), eps1=eps, stats=stats,
                        delta=delta, xout=xout)
        ctx.save_for_backward(xout, gamma, bias, stats)
        ctx.set_materialize_grads(False)
        return xout, y

    @staticmethod
    def backward(ctx, dxout, dy):
        xout, gamma, bias, stats = ctx.saved_tensors
        D = xout.shape[1]
        dg = zeros_f32(D, device=xout.device)
        db = zeros_f32(D, device=xout.device) if bias is not None else None
        if dy is None:      # only the residual stream was used downstream
            dx = dxout.contiguous()
            return dx, to_bf16(dx), dg, db, None
        dx = torch.empty_like(xout)
        dxb = torch.empty(xout.shape, dtype=bf16, device=xout.device)
        K.layernorm_bwd(dy.contiguous(), xout, gamma.detach(), stats, dx, dg, b1=None if bias is None else bias.detach(),
                        db1=db, dres=None if dxout is None else dxout.contiguous(), dx_bf16=dxb)
        return dx, dxb, dg, db, None


def add_layer_norm(x, delta, gamma, bias=None, eps=1e-5):
    return AddLayerNormFn.apply(x, delta, gamma, bias, eps)


class AddDeltaFn(torch.autograd.Function):
    """x (f32) + delta (bf16) -> f32: materialises a residual stream whose last delta is still pending"""

    @staticmethod
    def forward(ctx, x, delta):
        return K.add_bf16(x.contiguous(), delta.contiguous())

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return dy, to_bf16(dy)


class SelfAttentionFn(torch.autograd.Function):
    """unmasked multi-head self-attention on a fused [B*N, 3*H*dh] bf16 qkv buffer (decoder blocks)"""

    @staticmethod
    def forward(ctx, qkv, B, N, H, dh, scale):
        HD = H * dh
        o = torch.empty(B * N, HD, dtype=bf16, device=qkv.device)
        lse = torch.empty(B, H, N, dtype=f32, device=qkv.device)
        K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, B=B, H=H, Nq=N, Nk=N, dh=dh, scale=scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.dims = (B, N, H, dh, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        B, N, H, dh, scale = ctx.dims
        HD = H * dh
        do = to_bf16(do.contiguous())
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(B, H, N, dtype=f32, device=qkv.device)
        K.attn_bwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, do, dqkv[:, :HD], dqkv[:, HD:2 * HD],
                   dqkv[:, 2 * HD:], delta, B=B, H=H, Nq=N, Nk=N, dh=dh, scale=scale)
        return dqkv, None, None, None, None, None


class CrossAttentionFn(torch.autograd.Function):
    """unmasked multi-head cross-attention: q [B*Nq, H*dh], kv [B*Nk, 2*H*dh] (k | v), both bf16 token-major
    (decoder CrossAttention, multimae_utils.py:185-214)"""

    @staticmethod
    def forward(ctx, q, kv, B, Nq, Nk, H, dh, scale):
        HD = H * dh
        o = torch.empty(B * Nq, HD, dtype=bf16, device=q.device)
        lse = torch.empty(B, H, Nq, dtype=f32, device=q.device)
        K.attn_fwd(q, kv[:, :HD], kv[:, HD:], o, lse, B=B, H=H, Nq=Nq, Nk=Nk, dh=dh, scale=scale)
        ctx.save_for_backward(q, kv, o, lse)
        ctx.dims = (B, Nq, Nk, H, dh, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        B, Nq, Nk, H, dh, scale = ctx.dims
        HD = H * dh
        do = to_bf16(do.contiguous())
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        delta = torch.empty(B, H, Nq, dtype=f32, device=q.device)
        K.attn_bwd(q, kv[:, :HD], kv[:, HD:], o, lse, do, dq, dkv[:, :HD], dkv[:, HD:], delta, B=B, H=H, Nq=Nq, Nk=Nk,
                   dh=dh, scale=scale)
        return dq, dkv, None, None, None, None, None, None


class PoolAttnFn(torch.autograd.Function):
    """R queries (batch-invariant, [R, H*64] bf16) over planar kv rows with a dense mask; see kernels.pool_attn_fwd"""

    @staticmethod
    def forward(ctx, q, kv, mask_u8, mode, B, H, N, n_head, scale):
        ctx.q_dtype = q.dtype
        q = to_bf16(q)          # (an fp32 q keeps its GRADIENT in fp32; the forward value is the reference's bf16 Linear output)
        R = q.shape[0]
        out = torch.empty(B, R, H * 64, dtype=bf16, device=q.device)
        stat = torch.empty(B * R * H * 3, dtype=f32, device=q.device)
        K.pool_attn_fwd(q, kv, mask_u8, mode, out, stat, B=B, R=R, H=H, N=N, n_head=n_head, scale=scale, q_batched=False)
        ctx.save_for_backward(q, kv, mask_u8, mode, out, stat)
        ctx.dims = (B, R, H, N, n_head, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, kv, mask_u8, mode, out, stat = ctx.saved_tensors
        B, R, H, N, n_head, scale = ctx.dims
        dout = to_bf16(dout.contiguous())
        dq = zeros_f32(R, H * 64, device=q.device)
        dkv = torch.empty_like(kv)
        K.pool_attn_bwd(q, kv, mask_u8, mode, out, stat, dout, dq, dkv, B=B, R=R, H=H, N=N, n_head=n_head, scale=scale,
                        q_batched=False)
        return (dq if ctx.q_dtype == f32 else dq.to(bf16)), dkv, None, None, None, None, None, None, None


class UnpatchifyFn(torch.autograd.Function):
    """'b (nh nw) (c ph pw) -> b c (nh ph) (nw pw)' on bf16 (output_adapters_simple.py:183-186)"""

    @staticmethod
    def forward(ctx, tok, B, C, H, W, P):
        img = torch.empty(B, C, H, W, dtype=bf16, device=tok.device)
        K.unpatchify(tok.contiguous(), img, C, H, W, P)
        ctx.dims = (B, C, H, W, P, tok.shape)
        return img

    @staticmethod
    def backward(ctx, dimg):
        B, C, H, W, P, shape = ctx.dims
        dtok = torch.empty(shape, dtype=bf16, device=dimg.device)
        K.unpatchify(dtok, to_bf16(dimg.contiguous()), C, H, W, P, inverse=True)
        return dtok, None, None, None, None, None


class MaskedLossFn(torch.autograd.Function):
    """criterion.py:85-115 / :142-172 (norm_pix=False) fused; kind 0 = MSE, 1 = L1"""

    @staticmethod
    def forward(ctx, pred, target, mask, P, kind):
        pred = pred.contiguous()
        target = target.contiguous()
        B = pred.shape[0]
        work = torch.empty(2 * B + 2, dtype=f32, device=pred.device)
        loss = torch.empty(1, dtype=f32, device=pred.device)
        K.masked_loss_fwd(pred, target, mask, P, kind, work, loss)
        ctx.save_for_backward(pred, target, mask, work)
        ctx.pk = (P, kind)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        pred, target, mask, work = ctx.saved_tensors
        P, kind = ctx.pk
        dpred = torch.empty_like(pred)
        K.masked_loss_bwd(pred, target, mask, P, kind, work, dloss.reshape(1).to(f32).contiguous(), dpred)
        return dpred, None, None, None, None


class MaskedCEFn(torch.autograd.Function):
    """criterion.py:24-58 (label_smoothing 0) fused: logsumexp - target logit, patch mask, per-sample mean, batch nanmean"""

    @staticmethod
    def forward(ctx, logits, target, mask, P):
        logits = logits.contiguous()
        target = target.contiguous()
        B = logits.shape[0]
        work = torch.empty(2 * B + 2, dtype=f32, device=logits.device)
        loss = torch.empty(1, dtype=f32, device=logits.device)
        K.masked_ce_fwd(logits, target, mask, P, work, loss)
        ctx.save_for_backward(logits, target, mask, work)
        ctx.P = P
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        logits, target, mask, work = ctx.saved_tensors
        dlogits = torch.empty_like(logits)
        K.masked_ce_bwd(logits, target, mask, ctx.P, work, dloss.reshape(1).to(f32).contiguous(), dlogits)
        return dlogits, None, None, None


class DinoLossFn(torch.autograd.Function):
    """criterion.py:328-335 fused (forward + student gradient in one launch); the teacher gets no gradient"""

    @staticmethod
    def forward(ctx, student, teacher, student_temp, teacher_temp):
        s = student if student.stride(-1) == 1 else student.contiguous()
        t = teacher.detach()
        t = t if t.stride(-1) == 1 else t.contiguous()
        if t.dtype != s.dtype:
            s, t = s.float(), t.float()
        row_loss, ds = K.dino_loss(s, t, student_temp, teacher_temp)
        ctx.save_for_backward(ds)
        ctx.in_dtype = student.dtype
        return row_loss.mean()

    @staticmethod
    def backward(ctx, dloss):
        (ds,) = ctx.saved_tensors
        return (ds * dloss).to(ctx.in_dtype), None, None, None


class HardNegLossFn(torch.autograd.Function):
    """criterion.py:214-268 fused: forward + both input gradients in one call of three launches (kernels.hardneg_loss)"""

    @staticmethod
    def forward(ctx, out1, out2, tau_plus, beta, temperature, easy):
        a = out1.float()
        b = out2.float()
        a = a if a.stride(-1) == 1 else a.contiguous()
        b = b if b.stride(-1) == 1 else b.contiguous()
        loss, d1, d2 = K.hardneg_loss(a, b, tau_plus, beta, temperature, easy)
        ctx.save_for_backward(d1, d2)
        ctx.dtypes = (out1.dtype, out2.dtype)
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        d1, d2 = ctx.saved_tensors
        return (d1 * dloss).to(ctx.dtypes[0]), (d2 * dloss).to(ctx.dtypes[1]), None, None, None, None


class MatmulNTFn(torch.autograd.Function):
    """C = A . B^T with bf16 tensor-core operands and an fp32 result (what torch.mm does under the reference's
    autocast, e.g. the [2B, D] x [D, 2B] similarity of HardNegtive_loss, criterion.py:240).  A [M, K], B [N, K]."""

    @staticmethod
    def forward(ctx, a, b):
        M, Kd = a.shape
        N = b.shape[0]
        pad8 = lambda n: (n + 7) // 8 * 8        # TMA operands need 16-byte row pitches in every layout used below
        ab = K.cast_bf16(a.detach().float().contiguous(), rows_pad=pad8(M), cols_pad=pad8(Kd))
        bb = K.cast_bf16(b.detach().float().contiguous(), rows_pad=pad8(N), cols_pad=pad8(Kd))
        out = torch.empty(pad8(M), pad8(N), dtype=f32, device=a.device)
        K.gemm(ab, bb, out)
        ctx.save_for_backward(ab, bb)
        ctx.meta = (M, N, Kd, a.dtype, b.dtype)
        return out[:M, :N]

    @staticmethod
    def backward(ctx, dc):
        ab, bb = ctx.saved_tensors
        M, N, Kd, adt, bdt = ctx.meta
        dcb = K.cast_bf16(dc.float().contiguous(), rows_pad=ab.shape[0], cols_pad=bb.shape[0])
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty(ab.shape, dtype=f32, device=dc.device)
            K.gemm(dcb, bb, da, b_mn=True)                      # dA = dC . B
            da = da[:M, :Kd].to(adt)
        if ctx.needs_input_grad[1]:
            db = wgrad(dcb, ab)[:N, :Kd].to(bdt)                # dB = dC^T . A
        return da, db


# ------------------------------------------------------------------------------------------------
# token embedding: visible-patch im2col + projection GEMM written straight into the planar stream
# ------------------------------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """Builds the planar token stream X [B*nenc + B*F, D] (f32):
      head plane: for each modality m (in s1, s2, dem[, dnw] order) its n_m visible patches, projected with the
                  conv weight viewed as [D, C*P*P] + bias + sin-cos pos-emb (input_adapters.py:97-119
                  restricted to the visible patches, multimae.py:378-383)
      tail plane: fusion_tokens + fusion pos-emb broadcast over the batch (multimae.py:353-354).
    args: meta (dict), fusion_tokens [1,F,D], then per modality its tensors: kind "patch" (default) -> (image, proj_weight,
    proj_bias); kind "semseg" (meta["kinds"][m], SemSegInputAdapter input_adapters.py:209-328) -> (class map [B,H,W]
    int64, class_emb weight [NC, E], proj_weight [D, E, P, P], proj_bias).  A semantic map is embedded as ONE GEMM of the
    visible patches' one-hot rows [B*n, NC*P*P] against the table T[c, ph, pw, :] = W[:, :, ph, pw] . class_emb[c]: the
    embedding lookup and the Conv2d(k = s = P) of the reference folded together (the table is parameter-sized algebra,
    113 MFLOP at ViT-B, done with torch.einsum in fp32; every batch-sized contraction is mmf_gemm_bf16)."""

    @staticmethod
    def _arity(kind):
        return 4 if kind == "semseg" else 3

    @staticmethod
    def _forward_tokens(ctx, meta, fusion_tokens, mod_args):
        """All modalities through ONE im2col + ONE GEMM driven by the device token table meta["tok"] (global ids of the
        visible tokens in encoder order): no per-modality count ever reaches the host.  A [B*nenc, Kpad] holds each token's
        patch in its modality's column block (zeros elsewhere) and a one-hot modality flag; B = [W_0 | W_1 | .. | 0]; the
        epilogue adds (bias_m + pos_m)[patch] through the row map.  Zero columns add exact zeros to the fp32 accumulators,
        so every token's value equals the per-modality GEMM's."""
        B, D, P, Fn, nenc = meta["B"], meta["D"], meta["P"], meta["F"], meta["nenc"]
        dev = fusion_tokens.device
        Mh = B * nenc
        imgs = [mod_args[3 * m].contiguous() for m in range(len(meta["pos"]))]
        ws = [mod_args[3 * m + 1] for m in range(len(imgs))]
        bs = [mod_args[3 * m + 2] for m in range(len(imgs))]
        ks = [int(im.shape[1]) * P * P for im in imgs]
        col_off = [sum(ks[:m]) for m in range(len(ks))]
        ktot = sum(ks)
        kpad = ktot + 8                      # flag columns ktot .. ktot + M - 1 (16-byte row pitch kept)
        tok_off = [0]
        for pos in meta["pos"]:
            tok_off.append(tok_off[-1] + pos.shape[0])
        X = torch.empty(Mh + B * Fn, D, dtype=f32, device=dev)
        A = torch.empty(Mh, kpad, dtype=bf16, device=dev)
        if Mh > 0:
            K.im2col_tokens(imgs, meta["tok"], A, P, col_off, tok_off, ktot)
            table = torch.cat([pos + b.detach()[None, :] for pos, b in zip(meta["pos"], bs)], 0)     # [sum F_m, D] fp32
            K.gemm(A, w_hcat_bf16(ws, kpad), X, residual=table, res_row_map=meta["tok"], res_period=nenc)
        if Fn > 0:
            fus = (fusion_tokens.detach()[0] + meta["pos_fusion"]).contiguous()
            K.bcast_rows(fus, X[Mh:], B, Fn, D, Fn * D)
        ctx.meta = meta
        ctx.tokens_path = (A, ks, col_off, ktot, [w.shape for w in ws])
        return X

    @staticmethod
    def _backward_tokens(ctx, dX):
        meta = ctx.meta
        B, D, Fn, nenc = meta["B"], meta["D"], meta["F"], meta["nenc"]
        A, ks, col_off, ktot, wshapes = ctx.tokens_path
        Mh = B * nenc
        dX = dX.contiguous()
        grads: List[Optional[torch.Tensor]] = []
        if Mh > 0:
            dY = K.cast_bf16(dX[:Mh])
            dW = wgrad(dY, A)                                              # [D, kpad] fp32: weight blocks, then bias columns
            for m, (kk, c0, shp) in enumerate(zip(ks, col_off, wshapes)):
                grads += [None, dW[:, c0:c0 + kk].contiguous().view(shp), dW[:, ktot + m].contiguous()]
        else:
            for shp in wshapes:
                grads += [None, torch.zeros(shp, dtype=f32, device=dX.device), torch.zeros(D, dtype=f32, device=dX.device)]
        dfus = None
        if Fn > 0:
            dfus = torch.empty(1, Fn, D, dtype=f32, device=dX.device)
            K.reduce_batch(dX[Mh:], dfus, B, Fn, D, Fn * D)
        return (None, dfus) + tuple(grads)

    @staticmethod
    def forward(ctx, meta, fusion_tokens, *mod_args):
        ctx.tokens_path = None
        if meta.get("tok") is not None:
            return EmbedFn._forward_tokens(ctx, meta, fusion_tokens, mod_args)
        B, D, P, Fn, nenc = meta["B"], meta["D"], meta["P"], meta["F"], meta["nenc"]
        kinds = meta.get("kinds") or ["patch"] * len(meta["idx"])
        dev = fusion_tokens.device
        Mh = B * nenc
        X = torch.empty(Mh + B * Fn, D, dtype=f32, device=dev)
        saved = []
        off = 0
        ap = 0
        for m, idx in enumerate(meta["idx"]):
            args = mod_args[ap: ap + EmbedFn._arity(kinds[m])]
            ap += EmbedFn._arity(kinds[m])
            n = idx.numel()
            A = None
            if n > 0 and kinds[m] == "semseg":
                cls, emb, w, b = args
                NC = emb.shape[0]
                A = torch.zeros(B * n, NC * P * P, dtype=bf16, device=dev)
                K.onehot_im2col(cls.contiguous(), idx, A, P, NC)
                table = torch.einsum("deuv,ce->dcuv", w.detach().float(), emb.detach().float()).reshape(D, NC * P * P)
                K.gemm(A, K.cast_bf16(table.contiguous()), X[off:], bias=b.detach(), residual=meta["pos"][m], res_row_map=idx,
                       res_period=n, out_period=n, out_batch_rows=nenc)
            elif n > 0:
                img, w, b = args
                C = img.shape[1]
                A = torch.empty(B * n, C * P * P, dtype=bf16, device=dev)
                K.im2col_gather(img.contiguous(), idx, A, P)
                K.gemm(A, w_bf16(w), X[off:], bias=b.detach(), residual=meta["pos"][m], res_row_map=idx, res_period=n,
                       out_period=n, out_batch_rows=nenc)
            saved.append(A)
            off += n
        if Fn > 0:
            fus = (fusion_tokens.detach()[0] + meta["pos_fusion"]).contiguous()
            K.bcast_rows(fus, X[Mh:], B, Fn, D, Fn * D)
        ctx.meta = meta
        ctx.kinds = kinds
        ctx.saved_A = saved
        ctx.mod_params = []          # what the parameter gradients need besides A
        ap = 0
        for m in range(len(meta["idx"])):
            args = mod_args[ap: ap + EmbedFn._arity(kinds[m])]
            ap += EmbedFn._arity(kinds[m])
            ctx.mod_params.append((args[1].detach(), args[2].detach(), meta.get("padding_idx", {}).get(m)) if kinds[m] == "semseg"
                                  else (args[1].shape,))
        return X

    @staticmethod
    def backward(ctx, dX):
        if ctx.tokens_path is not None:
            return EmbedFn._backward_tokens(ctx, dX)
        meta = ctx.meta
        B, D, P, Fn, nenc = meta["B"], meta["D"], meta["P"], meta["F"], meta["nenc"]
        Mh = B * nenc
        dX = dX.contiguous()
        grads: List[Optional[torch.Tensor]] = []
        off = 0
        for m, idx in enumerate(meta["idx"]):
            n = idx.numel()
            semseg = ctx.kinds[m] == "semseg"
            if n == 0:
                if semseg:
                    emb, w, _ = ctx.mod_params[m]
                    grads += [None, torch.zeros_like(emb), torch.zeros_like(w), torch.zeros(D, dtype=f32, device=dX.device)]
                else:
                    grads += [None, torch.zeros(ctx.mod_params[m][0], dtype=f32, device=dX.device),
                              torch.zeros(D, dtype=f32, device=dX.device)]
                continue
            dY = torch.empty(B * n, D, dtype=bf16, device=dX.device)
            K.gather_rows(dX, dY, batch=B, n=n, d=D, src_batch_rows=nenc, row_off=off)
            db = K.colsum(dY, zeros_f32(D, device=dX.device))
            if semseg:
                emb, w, pad = ctx.mod_params[m]
                NC = emb.shape[0]
                dT = wgrad(dY, ctx.saved_A[m]).view(D, NC, P, P)                       # gradient of the table
                dw = torch.einsum("dcuv,ce->deuv", dT, emb.float())
                demb = torch.einsum("dcuv,deuv->ce", dT, w.float())
                if pad is not None:
                    demb[pad] = 0                                                        # nn.Embedding(padding_idx=...)
                grads += [None, demb, dw, db]
            else:
                dW = wgrad(dY, ctx.saved_A[m]).view(ctx.mod_params[m][0])
                grads += [None, dW, db]
            off += n
        dfus = None
        if Fn > 0:
            dfus = torch.empty(1, Fn, D, dtype=f32, device=dX.device)
            K.reduce_batch(dX[Mh:], dfus, B, Fn, D, Fn * D)
        return (None, dfus) + tuple(grads)


class PosEmbGradFn(torch.autograd.Function):
    """Gradient path of LEARNABLE positional embeddings (input_adapters.py:41-48, 76-87 with learnable_pos_emb=True or
    sincos_pos_emb=False).  EmbedFn adds the (detached) tables in its GEMM epilogue; this node is the identity on the token
    stream X and, in backward, returns d table[g] = sum over the batch of dX at the token whose global id is g (a batch
    reduction + an index_add through the token table) and d pos_fusion = the batch sum of the fusion rows' gradients.
    args: X [B*nenc + B*F, D], tok [nenc] int32, meta (B, nenc, F, sizes), table_all [sum F_m, D] or None, pos_fusion [F, D] or None"""

    @staticmethod
    def forward(ctx, X, tok, meta, table_all, pos_fusion):
        ctx.tok, ctx.meta = tok, meta
        ctx.rows = None if table_all is None else table_all.shape[0]
        return X.view_as(X)

    @staticmethod
    def backward(ctx, dX):
        B, nenc, Fn, D = ctx.meta["B"], ctx.meta["nenc"], ctx.meta["F"], dX.shape[1]
        dX = dX.contiguous()
        Mh = B * nenc
        dtab = dfus = None
        if ctx.needs_input_grad[3] and Mh > 0:
            dsum = torch.empty(1, nenc, D, dtype=f32, device=dX.device)
            K.reduce_batch(dX[:Mh], dsum, B, nenc, D, nenc * D)
            dtab = torch.zeros(ctx.rows, D, dtype=f32, device=dX.device).index_add_(0, ctx.tok.long(), dsum[0])
        if ctx.needs_input_grad[4] and Fn > 0:
            dfus = torch.empty(1, Fn, D, dtype=f32, device=dX.device)
            K.reduce_batch(dX[Mh:], dfus, B, Fn, D, Fn * D)
            dfus = dfus[0]
        return dX, None, None, dtab, dfus


# ------------------------------------------------------------------------------------------------
# the encoder stack
# ------------------------------------------------------------------------------------------------
ZB = 9  # tensors per (zorro or fusion) block: norm1.g, attn.norm.g, to_q.W, to_kv.W, to_out.W, norm2.g, mlp.0.g, mlp.1.W, mlp.3.W


def _ln2(x, g1, g2, x2=None, split=0, rows=None, delta=None, delta_row0=0, xout=None, y=None):
    """fused double LayerNorm of (x [+ delta]) -> bf16 (into `y` if given); see kernels.layernorm_fwd for the residual-add
    arguments"""
    rows = rows if rows is not None else x.shape[0]
    if y is None:
        y = torch.empty(rows, x.shape[1], dtype=bf16, device=x.device)
    st = torch.empty(rows, 4, dtype=f32, device=x.device)
    K.layernorm_fwd(x, g1, y, g2=g2, stats=st, x2=x2, x_split=split, rows=rows, delta=delta, delta_row0=delta_row0, xout=xout)
    return y, st


def _ffn_fwd(h, w1b, w2b, ipad):
    """GEGLU feed-forward without the residual add: returns (delta bf16 [rows, D], g, u).  The caller's next
    LayerNorm launch adds delta to the fp32 residual stream (reference: Linear output bf16, sum fp32)."""
    rows = h.shape[0]
    g = torch.empty(rows, ipad, dtype=bf16, device=h.device)
    u = torch.empty(rows, 2 * ipad, dtype=bf16, device=h.device)
    K.gemm(h, w1b, g, act=2, out2=u)
    delta = torch.empty(rows, w2b.shape[0], dtype=bf16, device=h.device)
    K.gemm(g, w2b, delta)
    return delta, g, u


def _unpad_w1(dw1, I, ipad):
    if I == ipad:
        return dw1
    return torch.cat([dw1[:I], dw1[ipad:ipad + I]], 0)


class _ZeroArena:
    """One zero-filled fp32 buffer per encoder layer for every gradient the layer accumulates into (split-K weight
    gradients, LayerNorm gammas): a single memset instead of a dozen small fill launches per layer."""

    def __init__(self, numel: int, device):
        self.buf = torch.zeros(numel, dtype=f32, device=device)
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.buf.numel():          # (bound was too tight: fall back to a private buffer)
            return torch.zeros(*shape, dtype=f32, device=self.buf.device)
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + 3) // 4 * 4                  # keep every view 16-byte aligned
        return v


def _layer_hook(meta, grads, slots, arena=None):
    """hand one encoder layer's parameter gradients to the data-parallel reducer while the backward continues
    (training.GradAllReduce).  Gradients that live in the layer's zero arena are reduced IN PLACE as one contiguous
    buffer (`grad_hook_inplace`: no packing copy); anything else goes through `grad_hook`, which returns the tensors
    autograd should install instead."""
    hook, inplace = meta.get("grad_hook"), meta.get("grad_hook_inplace")
    if hook is None and inplace is None:
        return
    slots = [j for j in slots if grads[j] is not None]
    rest = slots
    if inplace is not None and arena is not None and arena.off > 0:
        lo = arena.buf.data_ptr()
        hi = lo + arena.off * 4
        inside = [j for j in slots if grads[j].is_contiguous() and lo <= grads[j].data_ptr() and
                  grads[j].data_ptr() + grads[j].numel() * 4 <= hi]
        if inside:
            inplace(arena.buf[:arena.off], [grads[j] for j in inside])
            rest = [j for j in slots if j not in set(inside)]
    if rest and hook is not None:
        out = hook([grads[j].contiguous() for j in rest])
        for j, g in zip(rest, out):
            grads[j] = g


class EncoderStackFn(torch.autograd.Function):
    """All encoder layers as one node.  meta: dims + the per-step mask structures (device int32 `seg`
    table, `slotmap`), fusion flag.  Flat params: [mask_embedding] (fusion variant) then per layer the
    9 fusion-block tensors (fusion variant) followed by the 9 zorro-block tensors."""

    @staticmethod
    def forward(ctx, meta, X, *params):
        B, D, H, Fn, nenc, fusion = meta["B"], meta["D"], meta["H"], meta["F"], meta["nenc"], meta["fusion"]
        depth, I = meta["depth"], meta["I"]
        ipad = _pad64(I)
        HD = H * 64
        Mh, Mf = B * nenc, B * Fn
        Mt = Mh + Mf
        N = nenc + Fn
        seg, nseg, scale = meta["seg"], meta["nseg"], 0.125
        per_layer = ZB * (2 if fusion else 1)
        base = 1 if fusion else 0
        me = params[0].detach()[0] if fusion else None
        saved = []
        # The fp32 residual stream is S (+ pend): `pend` is the bf16 output of the last sub-layer GEMM that has not
        # been added yet; the next LayerNorm launch adds it and writes the materialised stream (one pass, no fp32
        # read-modify-write in a GEMM epilogue).
        S = X.contiguous()
        pend = None
        # optional pyramid taps (downstream ViTBaseline, multimae_big_imcomplete.py:651-653): the fusion tokens after the
        # listed blocks are returned as extra outputs, in ascending block order
        taps = sorted(meta.get("taps") or [])
        tap_out = []
        for i in range(depth):
            lp = [p.detach() for p in params[base + i * per_layer: base + (i + 1) * per_layer]]
            rec = {}
            Xf2 = None
            Xin = S if pend is None else torch.empty(Mt, D, dtype=f32, device=S.device)
            if fusion:
                fn1, fan, fwq, fwkv, fwo, fn2, fm0, fw1, fw2 = lp[:ZB]
                # the normalised mask-embedding rows ride at the end of the token rows: ONE k/v projection (and, in the
                # backward, one dgrad and one wgrad) per layer instead of a second, 196-row launch of each
                hk_all = torch.empty(Mt + Fn, D, dtype=bf16, device=S.device)
                hk, hm = hk_all[:Mt], hk_all[Mt:]
                _, stA = _ln2(S, fn1, fan, delta=pend, xout=Xin if pend is not None else None, y=hk)
                _, stM = _ln2(me.contiguous(), fn1, fan, y=hm)
                wkv = w_bf16(params[base + i * per_layer + 3])
                kv_all = torch.empty(Mt + Fn, 2 * HD, dtype=bf16, device=S.device)
                K.gemm(hk_all, wkv, kv_all)
                kv, kvm = kv_all[:Mt], kv_all[Mt:]
                q = torch.empty(Mf, HD, dtype=bf16, device=S.device)
                K.gemm(hk[Mh:], w_bf16(params[base + i * per_layer + 2]), q)
                a = torch.empty(Mf, HD, dtype=bf16, device=S.device)
                K.slot_attn_fwd(q, kv, kvm, meta["slotmap"], seg, a, None, B=B, F=Fn, H=H, S=nseg, n_head=nenc, scale=scale)
                dA = torch.empty(Mf, D, dtype=bf16, device=S.device)
                K.gemm(a, w_bf16(params[base + i * per_layer + 4]), dA)
                Xf1 = torch.empty(Mf, D, dtype=f32, device=S.device)
                h2, stB = _ln2(Xin[Mh:], fn2, fm0, delta=dA, xout=Xf1)            # Xf1 = fusion tokens + attention
                dF, g, u = _ffn_fwd(h2, w_geglu_bf16(params[base + i * per_layer + 7], ipad),
                                    w_bf16(params[base + i * per_layer + 8], cols_pad=ipad), ipad)
                Xf2 = torch.empty(Mf, D, dtype=f32, device=S.device)
                rec.update(hk=hk, hk_all=hk_all, stA=stA, kv=kv, q=q, hm=hm, stM=stM, kvm=kvm, a=a, Xf1=Xf1, h2=h2, stB=stB, g=g, u=u,
                           Xf2=Xf2)
            zo = base + i * per_layer + (ZB if fusion else 0)
            n1, an, wq, wkv_, wo, n2, m0, w1, w2 = [p.detach() for p in params[zo: zo + ZB]]
            if fusion:   # rows >= Mh: Xf2 = Xf1 + ffn (written by this launch), rows < Mh: the block input
                h1, st1 = _ln2(Xin, n1, an, x2=Xf1, split=Mh, rows=Mt, delta=dF, delta_row0=Mh, xout=Xf2)
            else:
                h1, st1 = _ln2(S, n1, an, delta=pend, xout=Xin if pend is not None else None)
            rec["X"] = Xin
            qkv = torch.empty(Mt, 3 * HD, dtype=bf16, device=S.device)
            K.gemm(h1, w_cat_bf16([params[zo + 2], params[zo + 3]]), qkv)
            o = torch.empty(Mt, HD, dtype=bf16, device=S.device)
            lse = torch.empty(B, H, N, dtype=f32, device=S.device)
            K.attn_fwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], o, lse, B=B, H=H, Nq=N, Nk=N, dh=64, scale=scale,
                       n_head_q=nenc, n_head_k=nenc, seg=seg, nseg=nseg)
            dO = torch.empty(Mt, D, dtype=bf16, device=S.device)
            K.gemm(o, w_bf16(params[zo + 4]), dO)
            X1 = torch.empty(Mt, D, dtype=f32, device=S.device)
            h2z, st2 = _ln2(Xin, n2, m0, x2=Xf2, split=Mh if fusion else 0, rows=Mt, delta=dO, xout=X1)   # X1 = x + attn
            dZ, gz, uz = _ffn_fwd(h2z, w_geglu_bf16(params[zo + 7], ipad), w_bf16(params[zo + 8], cols_pad=ipad), ipad)
            rec.update(h1=h1, st1=st1, qkv=qkv, o=o, lse=lse, X1=X1, h2z=h2z, st2=st2, gz=gz, uz=uz)
            saved.append(rec)
            S, pend = X1, dZ
            if i in taps and i != depth - 1:   # fusion tokens after block i (stream + its pending delta), fp32 [B*F, D]
                tap_out.append(K.add_bf16(S[Mh:], pend[Mh:]))
        X = K.add_bf16(S, pend) if pend is not None else S
        if (depth - 1) in taps:
            tap_out.append(X[Mh:].clone())   # (an output may not be a view of another output)
        ctx.meta = meta
        ctx.saved = saved
        ctx.params = params
        ctx.taps = taps
        if taps:
            return (X,) + tuple(tap_out)
        return X

    @staticmethod
    def backward(ctx, dX, *dtaps):
        if ctx.saved is None:
            raise RuntimeError("EncoderStackFn.backward ran twice: the encoder stack frees each layer's activations as its "
                               "backward consumes them (retain_graph is not supported through this node)")
        meta, saved, params = ctx.meta, ctx.saved, ctx.params
        tap_grad = {i: g for i, g in zip(ctx.taps, dtaps) if g is not None}
        B, D, H, Fn, nenc, fusion = meta["B"], meta["D"], meta["H"], meta["F"], meta["nenc"], meta["fusion"]
        depth, I = meta["depth"], meta["I"]
        ipad = _pad64(I)
        HD = H * 64
        Mh, Mf = B * nenc, B * Fn
        Mt = Mh + Mf
        N = nenc + Fn
        seg, nseg, scale = meta["seg"], meta["nseg"], 0.125
        per_layer = ZB * (2 if fusion else 1)
        base = 1 if fusion else 0
        dev = dX.device
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        dme = torch.zeros(Fn, D, dtype=f32, device=dev) if fusion else None
        dX = dX.contiguous()
        if tap_grad:
            dX = dX.clone()     # tap gradients are accumulated into it in place below
        # per layer: every parameter gradient (weights at their padded GEGLU width) + the mask-embedding k/v gradient
        layer_elems = (2 if fusion else 1) * (4 * D + 4 * HD * D + 3 * ipad * D + 64) + (Fn * 2 * HD if fusion else 0)
        arena = None
        zeros = lambda n: arena.take(n)

        def ffn_bwd(dXo_b, g, u, h, w1_param, w2_param, rows):
            """returns dh (bf16 [rows, D]), dW1, dW2"""
            w2b = w_bf16(w2_param, cols_pad=ipad)
            # dg = dXo . W2 stays in TMEM: the GEMM's epilogue turns it into du = [dvalue | dgate] against the saved u
            du = torch.empty_like(u)
            K.gemm(dXo_b, w2b, du, b_mn=True, act=3, out2=u)
            dW2 = wgrad(dXo_b, g, out=arena.take(D, ipad))[:, :I]
            w1b = w_geglu_bf16(w1_param, ipad)
            dh = torch.empty(rows, D, dtype=bf16, device=dev)
            K.gemm(du, w1b, dh, b_mn=True)
            dW1 = _unpad_w1(wgrad(du, h, out=arena.take(2 * ipad, D)), I, ipad)
            return dh, dW1, dW2.contiguous() if I != ipad else dW2

        dXb_next = None
        for i in reversed(range(depth)):
            rec = saved[i]
            saved[i] = None
            arena = _ZeroArena(layer_elems, dev)
            if i in tap_grad:   # dX is the gradient w.r.t. block i's output: the tap is its fusion plane
                K.add_inplace(dX[Mh:], tap_grad[i].contiguous().float().view(Mf, D))
                dXb_next = None                         # the bf16 copy emitted by the previous LayerNorm backward is stale
            X = rec["X"]
            Xf2 = rec.get("Xf2")
            zo = base + i * per_layer + (ZB if fusion else 0)
            n1, an, n2, m0 = [params[zo + j].detach() for j in (0, 1, 5, 6)]
            # ---------------- zorro block backward ----------------
            dX2b = dXb_next if dXb_next is not None else K.cast_bf16(dX)   # bf16 copy emitted by the previous LN backward
            dh2z, dW1, dW2 = ffn_bwd(dX2b, rec["gz"], rec["uz"], rec["h2z"], params[zo + 7], params[zo + 8], Mt)
            grads[zo + 7], grads[zo + 8] = dW1, dW2
            dX1 = torch.empty(Mt, D, dtype=f32, device=dev)
            dX1b = torch.empty(Mt, D, dtype=bf16, device=dev)
            dn2, dm0 = zeros(D), zeros(D)
            K.layernorm_bwd(dh2z, rec["X1"], n2, rec["st2"], dX1, dn2, g2=m0, dres=dX, dx_bf16=dX1b, dg2=dm0)
            grads[zo + 5], grads[zo + 6] = dn2, dm0
            do = torch.empty(Mt, HD, dtype=bf16, device=dev)
            K.gemm(dX1b, w_bf16(params[zo + 4]), do, b_mn=True)
            grads[zo + 4] = wgrad(dX1b, rec["o"], out=arena.take(D, HD))
            qkv = rec["qkv"]
            dqkv = torch.empty_like(qkv)
            delta = torch.empty(B, H, N, dtype=f32, device=dev)
            K.attn_bwd(qkv[:, :HD], qkv[:, HD:2 * HD], qkv[:, 2 * HD:], rec["o"], rec["lse"], do, dqkv[:, :HD],
                       dqkv[:, HD:2 * HD], dqkv[:, 2 * HD:], delta, B=B, H=H, Nq=N, Nk=N, dh=64, scale=scale,
                       n_head_q=nenc, n_head_k=nenc, seg=seg, nseg=nseg)
            dh1 = torch.empty(Mt, D, dtype=bf16, device=dev)
            K.gemm(dqkv, w_cat_bf16([params[zo + 2], params[zo + 3]]), dh1, b_mn=True)
            dWqkv = wgrad(dqkv, rec["h1"], out=arena.take(3 * HD, D))
            grads[zo + 2], grads[zo + 3] = dWqkv[:HD], dWqkv[HD:]
            dZ = torch.empty(Mt, D, dtype=f32, device=dev)
            dZb = torch.empty(Mt, D, dtype=bf16, device=dev) if (fusion or i > 0) else None
            dn1, dan = zeros(D), zeros(D)
            K.layernorm_bwd(dh1, X, n1, rec["st1"], dZ, dn1, g2=an, dres=dX1, dx_bf16=dZb, dg2=dan, x2=Xf2,
                            x_split=Mh if fusion else 0, rows=Mt)
            grads[zo], grads[zo + 1] = dn1, dan
            if not fusion:
                dX, dXb_next = dZ, dZb
                _layer_hook(meta, grads, range(zo, zo + ZB), arena)
                continue
            # ---------------- fusion block backward (upstream: dZ[Mh:] = grad wrt Xf2) ----------------
            fo = base + i * per_layer
            fn1, fan, fn2, fm0 = [params[fo + j].detach() for j in (0, 1, 5, 6)]
            dh2, dW1, dW2 = ffn_bwd(dZb[Mh:], rec["g"], rec["u"], rec["h2"], params[fo + 7], params[fo + 8], Mf)
            grads[fo + 7], grads[fo + 8] = dW1, dW2
            dXf1b = torch.empty(Mf, D, dtype=bf16, device=dev)
            dfn2, dfm0 = zeros(D), zeros(D)
            # in place: dZ[Mh:] (= dXf2) becomes dXf1, which is also the gradient of the fusion residual input
            K.layernorm_bwd(dh2, rec["Xf1"], fn2, rec["stB"], dZ[Mh:], dfn2, g2=fm0, dres=dZ[Mh:], dx_bf16=dXf1b, dg2=dfm0)
            grads[fo + 5], grads[fo + 6] = dfn2, dfm0
            da = torch.empty(Mf, HD, dtype=bf16, device=dev)
            K.gemm(dXf1b, w_bf16(params[fo + 4]), da, b_mn=True)
            grads[fo + 4] = wgrad(dXf1b, rec["a"], out=arena.take(D, HD))
            # dkv [Mt, 2HD] and dq [Mf, HD] share one buffer, [dk | dv | dq] per row (the dq columns of the modality rows
            # stay unused): the fusion rows' dgrad is then ONE GEMM over K = 3HD against [Wkv; Wq] instead of a second,
            # read-modify-write GEMM into the same rows
            # rows [0, Mt): tokens; rows [Mt, Mt + Fn): the mask-embedding rows (batch-invariant keys / values), so that
            # their k/v gradient shares the token rows' dgrad and wgrad launches
            dkvq = torch.empty(Mt + Fn, 3 * HD, dtype=bf16, device=dev)
            dkv, dq = dkvq[:Mt, :2 * HD], dkvq[Mh:Mt, 2 * HD:]
            dkvm = arena.take(Fn, 2 * HD)
            K.slot_attn_bwd(rec["q"], rec["kv"], rec["kvm"], meta["slotmap"], seg, da, dq, dkv, dkvm, B=B, F=Fn, H=H, S=nseg,
                            n_head=nenc, scale=scale)
            K.cast_bf16(dkvm, dkvq[Mt:, :2 * HD])
            wkvb = w_bf16(params[fo + 3])
            dhk_all = torch.empty(Mt + Fn, D, dtype=bf16, device=dev)
            dhk, dhm = dhk_all[:Mt], dhk_all[Mt:]
            K.gemm(dkvq[:Mh, :2 * HD], wkvb, dhk_all[:Mh], b_mn=True)
            K.gemm(dkvq[Mh:Mt], w_cat_bf16([params[fo + 3], params[fo + 2]]), dhk_all[Mh:Mt], b_mn=True)
            K.gemm(dkvq[Mt:, :2 * HD], wkvb, dhm, b_mn=True)
            grads[fo + 3] = wgrad(dkvq[:, :2 * HD], rec["hk_all"], out=arena.take(2 * HD, D))
            grads[fo + 2] = wgrad(dq, rec["hk"][Mh:], out=arena.take(HD, D))
            dfn1, dfan = zeros(D), zeros(D)
            dme_i = torch.empty(Fn, D, dtype=f32, device=dev)
            K.layernorm_bwd(dhm, params[0].detach()[0].contiguous(), fn1, rec["stM"], dme_i, dfn1, g2=fan, dg2=dfan)
            K.add_inplace(dme, dme_i)
            dXin = torch.empty(Mt, D, dtype=f32, device=dev)
            dXb_next = torch.empty(Mt, D, dtype=bf16, device=dev) if i > 0 else None
            K.layernorm_bwd(dhk, X, fn1, rec["stA"], dXin, dfn1, g2=fan, dres=dZ, dx_bf16=dXb_next, dg2=dfan)
            grads[fo], grads[fo + 1] = dfn1, dfan
            dX = dXin
            _layer_hook(meta, grads, range(fo, fo + 2 * ZB), arena)
        if fusion:
            grads[0] = dme.view(1, Fn, D)
        ctx.saved = None
        return (None, dX) + tuple(grads)
