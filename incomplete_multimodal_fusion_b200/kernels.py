"""Thin launch wrappers: torch tensors in, raw pointers + sizes across the C ABI (include/mmf_b200.h).

torch is used here only for device memory and the current stream.  None of these functions is
autograd-aware; `functions.py` composes them into `torch.autograd.Function`s.  Every wrapper raises
if a tensor is not on a CUDA device: there is no CPU path.
"""
import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import AttnArgs, GemmArgs, PoolAttnArgs, SlotAttnArgs, check

bf16 = torch.bfloat16
f32 = torch.float32


def _L():
    return _lib.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("incomplete_multimodal_fusion_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
    return t.data_ptr()


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor, got shape %s strides %s" % (tuple(t.shape), t.stride())
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# Optional per-launch timing of the GEMM kernel (bench.py's roofline leg): a list of
# (start_event, end_event, flops, kind) appended around every mmf_gemm_bf16 launch while enabled.
GEMM_TIMING = None


def enable_gemm_timing(flag: bool = True):
    global GEMM_TIMING
    GEMM_TIMING = [] if flag else None
    return GEMM_TIMING


# Same idea for the non-GEMM kernels bench.py reports a roofline for: name -> [(start, end, meta)]
KERNEL_TIMING = None


def enable_kernel_timing(flag: bool = True):
    global KERNEL_TIMING
    KERNEL_TIMING = {} if flag else None
    return KERNEL_TIMING


class _Timed:
    """with _Timed(name, meta): launch  -- records CUDA events on the launch stream when KERNEL_TIMING is enabled"""

    def __init__(self, name, meta):
        self.name, self.meta = name, meta

    def __enter__(self):
        if KERNEL_TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if KERNEL_TIMING is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            KERNEL_TIMING.setdefault(self.name, []).append((self.e0, e1, self.meta))
        return False


def launch_count() -> int:
    return int(_L().mmf_launch_count())


def set_gemm_reserved_sms(n: int):
    """leave n SMs out of the persistent GEMM grids (data-parallel runs: room for the concurrent NCCL all-reduce)"""
    _L().mmf_set_gemm_reserved_sms(int(n))


def reset_launch_count():
    _L().mmf_reset_launch_count()


# ------------------------------------------------------------------------------------------------
def gemm(a, b, out, *, a_mn=False, b_mn=False, bias=None, act=0, residual=None, residual2=None, res_split=0,
         res_row_map=None, res_period=0, out_period=0, out_batch_rows=0, split_k=1, alpha=1.0, out2=None, block_n=0,
         accumulate=False, M=None, N=None, K=None):
    """out[M,N] = epilogue(alpha * A . B^T).  a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn),
    both bf16 row-major 2-D.  act=2 (GEGLU): b is [2*Ipad, K] and N = Ipad.  act=3 (GEGLU backward fused into the
    dgrad GEMM): out2 = saved [value | gate] (input), out = [dvalue | dgate], both [M, 2*Ipad]; N = Ipad."""
    assert a.dtype == bf16 and b.dtype == bf16
    if M is None:
        M = a.shape[1] if a_mn else a.shape[0]
    if K is None:
        K = a.shape[0] if a_mn else a.shape[1]
    if N is None:
        N = b.shape[1] if b_mn else b.shape[0]
        if act == 2:
            N //= 2
    kb = b.shape[0] if b_mn else b.shape[1]
    assert kb == K, "GEMM inner dimensions differ: %d vs %d" % (K, kb)
    assert out.dtype in (bf16, f32)
    args = GemmArgs()
    args.a, args.b, args.out = _p(a), _p(b), _p(out)
    args.bias, args.residual, args.residual2 = _p(bias), _p(residual), _p(residual2)
    args.res_split = res_split
    args.res_row_map = _p(res_row_map)
    args.M, args.N, args.K = M, N, K
    args.lda, args.ldb, args.ldo = _ld(a), _ld(b), _ld(out)
    args.ldr = _ld(residual) if residual is not None else 0
    if residual is not None:
        assert residual.dtype == f32
    if residual2 is not None:
        assert residual2.dtype == f32 and _ld(residual2) == args.ldr
    if bias is not None:
        assert bias.dtype == f32 and bias.is_contiguous()
    if res_row_map is not None:
        assert res_row_map.dtype == torch.int32
    args.a_mn, args.b_mn = int(a_mn), int(b_mn)
    args.out_f32 = int(out.dtype == f32)
    args.act, args.split_k = act, split_k
    args.res_period, args.out_period, args.out_batch_rows = res_period, out_period, out_batch_rows
    args.block_n = block_n
    args.alpha = alpha
    args.accumulate = int(accumulate)
    args.out2 = _p(out2)
    args.ldo2 = _ld(out2) if out2 is not None else 0
    if out2 is not None:
        assert out2.dtype == bf16
    if GEMM_TIMING is None:
        check(_L().mmf_gemm_bf16(C.byref(args), _stream()), "mmf_gemm_bf16")
        return out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(_L().mmf_gemm_bf16(C.byref(args), _stream()), "mmf_gemm_bf16")
    e1.record()
    n_eff = 2 * N if act == 2 else N
    kind = "wgrad" if (a_mn and b_mn) else ("dgrad" if b_mn else "fwd")
    GEMM_TIMING.append((e0, e1, 2.0 * M * n_eff * K, kind, (kind + {2: "_geglu", 3: "_geglubwd"}.get(act, ""), M, n_eff, K)))
    return out


def layernorm_fwd(x, g1, y, *, b1=None, eps1=1e-5, g2=None, eps2=1e-5, stats=None, x2=None, x_split=0, rows=None,
                  delta=None, delta_row0=0, xout=None):
    """y = LN2(LN1(x')) with x' = x (+ delta for rows >= delta_row0, in which case x' of those rows is written to xout)"""
    assert x.dtype == f32 and y.dtype in (bf16, f32)
    rows = rows if rows is not None else x.shape[0]
    D = x.shape[1]
    if delta is not None:
        assert delta.dtype == bf16 and xout is not None and xout.dtype == f32
        assert delta.shape[0] >= rows - delta_row0 and xout.shape[0] >= rows - delta_row0
    # algorithmic bytes: read x (4), write y (2 or 4), + delta rows: read delta (2), write xout (4)
    nbytes = rows * D * (4 + y.element_size()) + (max(rows - delta_row0, 0) * D * 6 if delta is not None else 0)
    with _Timed("ln_fwd", nbytes):
        check(_L().mmf_layernorm_fwd(_p(x), _p(x2), x_split, rows, D, _ld(x), _p(g1), _p(b1), eps1, _p(g2), eps2, _p(y), _ld(y),
                                     int(y.dtype == f32), _p(stats), _p(delta), delta_row0, _ld(delta) if delta is not None else 0,
                                     _p(xout), _ld(xout) if xout is not None else 0, _stream()), "mmf_layernorm_fwd")
    return y


def layernorm_bwd(dy, x, g1, stats, dx, dg1, *, b1=None, g2=None, dres=None, dx_bf16=None, db1=None, dg2=None, x2=None,
                  x_split=0, rows=None):
    assert dy.dtype in (bf16, f32) and x.dtype == f32 and dx.dtype == f32
    rows = rows if rows is not None else dy.shape[0]
    D = dy.shape[1]
    # algorithmic bytes: read dy, x (4) (+ dres 4), write dx (4) (+ bf16 copy 2)
    nbytes = rows * D * (dy.element_size() + 8 + (4 if dres is not None else 0) + (2 if dx_bf16 is not None else 0))
    with _Timed("ln_bwd", nbytes):
        check(_L().mmf_layernorm_bwd(_p(dy), _ld(dy), int(dy.dtype == f32), _p(x), _p(x2), x_split, rows, D, _ld(x), _p(g1), _p(b1),
                                     _p(g2), _p(stats), _p(dres), _ld(dres) if dres is not None else 0, _p(dx), _ld(dx),
                                     _p(dx_bf16), _ld(dx_bf16) if dx_bf16 is not None else 0, _p(dg1), _p(db1), _p(dg2), _stream()),
              "mmf_layernorm_bwd")
    return dx


def _attn_args(q, k, v, o, lse, B, H, Nq, Nk, dh, scale, n_head_q, n_head_k, seg, nseg):
    a = AttnArgs()
    a.q, a.k, a.v, a.o, a.lse = _p(q), _p(k), _p(v), _p(o), _p(lse)
    a.ldq, a.ldk, a.ldv, a.ldo = q.stride(0), k.stride(0), v.stride(0), o.stride(0)
    a.B, a.H, a.Nq, a.Nk, a.dh = B, H, Nq, Nk, dh
    a.n_head_q = Nq if n_head_q is None else n_head_q
    a.n_tail_q = Nq - a.n_head_q
    a.n_head_k = Nk if n_head_k is None else n_head_k
    a.n_tail_k = Nk - a.n_head_k
    a.scale = scale
    a.seg = _p(seg)
    a.nseg = nseg
    return a


def attn_fwd(q, k, v, o, lse, *, B, H, Nq, Nk, dh, scale, n_head_q=None, n_head_k=None, seg=None, nseg=0):
    """q/k/v/o: bf16 2-D token-major views (column slices of a fused qkv buffer are fine)."""
    a = _attn_args(q, k, v, o, lse, B, H, Nq, Nk, dh, scale, n_head_q, n_head_k, seg, nseg)
    with _Timed("attn_fwd", (B, H, Nq, Nk, dh, seg, nseg)):
        check(_L().mmf_attn_fwd(C.byref(a), _stream()), "mmf_attn_fwd")
    return o


def attn_bwd(q, k, v, o, lse, d_o, dq, dk, dv, delta, *, B, H, Nq, Nk, dh, scale, n_head_q=None, n_head_k=None, seg=None,
             nseg=0):
    a = _attn_args(q, k, v, o, lse, B, H, Nq, Nk, dh, scale, n_head_q, n_head_k, seg, nseg)
    a.d_o, a.lddo, a.delta = _p(d_o), d_o.stride(0), _p(delta)
    a.dq, a.dk, a.dv = _p(dq), _p(dk), _p(dv)
    a.lddq, a.lddk, a.lddv = dq.stride(0), dk.stride(0), dv.stride(0)
    with _Timed("attn_bwd", (B, H, Nq, Nk, dh, seg, nseg)):
        check(_L().mmf_attn_bwd(C.byref(a), _stream()), "mmf_attn_bwd")


def _slot_args(q, kv_tok, kv_me, slotmap, seg, B, F, H, S, n_head, scale):
    a = SlotAttnArgs()
    a.q, a.kv_tok, a.kv_me, a.slotmap, a.seg = _p(q), _p(kv_tok), _p(kv_me), _p(slotmap), _p(seg)
    a.ldq, a.ldkv, a.ldme = q.stride(0), kv_tok.stride(0), kv_me.stride(0)
    a.B, a.F, a.H, a.S, a.dh, a.n_head, a.scale = B, F, H, S, 64, n_head, scale
    return a


def slot_attn_fwd(q, kv_tok, kv_me, slotmap, seg, out, probs, *, B, F, H, S, n_head, scale):
    a = _slot_args(q, kv_tok, kv_me, slotmap, seg, B, F, H, S, n_head, scale)
    a.out, a.ldo, a.probs = _p(out), out.stride(0), _p(probs)
    check(_L().mmf_slot_attn_fwd(C.byref(a), _stream()), "mmf_slot_attn_fwd")
    return out


def slot_attn_bwd(q, kv_tok, kv_me, slotmap, seg, dout, dq, dkv_tok, dkv_me, *, B, F, H, S, n_head, scale):
    a = _slot_args(q, kv_tok, kv_me, slotmap, seg, B, F, H, S, n_head, scale)
    a.dout, a.dq, a.dkv_tok, a.dkv_me = _p(dout), _p(dq), _p(dkv_tok), _p(dkv_me)
    scratch = torch.empty(B * F * H * (S - 1) * 2, dtype=f32, device=q.device)   # (ds, w) per sample / position / head / slot
    a.me_scratch = _p(scratch)
    a.lddout, a.lddq, a.lddkv, a.lddme = dout.stride(0), dq.stride(0), dkv_tok.stride(0), dkv_me.stride(0)
    check(_L().mmf_slot_attn_bwd(C.byref(a), _stream()), "mmf_slot_attn_bwd")


def _pool_args(q, kv, mask, mode, out, stat, B, R, H, N, n_head, scale, q_batched):
    a = PoolAttnArgs()
    a.q, a.kv, a.mask, a.mode, a.out, a.stat = _p(q), _p(kv), _p(mask), _p(mode), _p(out), _p(stat)
    a.q_bstride = R * H * 64 if q_batched else 0
    a.ldkv = kv.stride(0)
    a.B, a.R, a.H, a.N, a.dh, a.n_head, a.n_tail, a.scale = B, R, H, N, 64, n_head, N - n_head, scale
    return a


def pool_attn_fwd(q, kv, mask, mode, out, stat, *, B, R, H, N, n_head, scale, q_batched):
    a = _pool_args(q, kv, mask, mode, out, stat, B, R, H, N, n_head, scale, q_batched)
    check(_L().mmf_pool_attn_fwd(C.byref(a), _stream()), "mmf_pool_attn_fwd")
    return out


def pool_attn_bwd(q, kv, mask, mode, out, stat, dout, dq, dkv, *, B, R, H, N, n_head, scale, q_batched):
    a = _pool_args(q, kv, mask, mode, out, stat, B, R, H, N, n_head, scale, q_batched)
    a.dout, a.dq, a.dkv = _p(dout), _p(dq), _p(dkv)
    a.dq_bstride = R * H * 64 if q_batched else 0
    a.lddkv = dkv.stride(0)
    check(_L().mmf_pool_attn_bwd(C.byref(a), _stream()), "mmf_pool_attn_bwd")


def masked_loss_fwd(pred, target, mask, P, kind, work, loss):
    B, Cc, H, W = pred.shape
    assert pred.is_contiguous() and target.is_contiguous() and target.dtype == f32
    mb = mask.stride(0) if mask is not None else 0
    if mask is not None:
        assert mask.dtype == torch.int64 and mask.stride(1) == 1
    check(_L().mmf_masked_loss_fwd(_p(pred), int(pred.dtype == f32), _p(target), _p(mask), mb, B, Cc, H, W, P, kind, _p(work),
                                   _p(loss), _stream()), "mmf_masked_loss_fwd")


def masked_loss_bwd(pred, target, mask, P, kind, work, dloss, dpred):
    B, Cc, H, W = pred.shape
    mb = mask.stride(0) if mask is not None else 0
    check(_L().mmf_masked_loss_bwd(_p(pred), int(pred.dtype == f32), _p(target), _p(mask), mb, B, Cc, H, W, P, kind, _p(work),
                                   _p(dloss), _p(dpred), _stream()), "mmf_masked_loss_bwd")


def masked_ce_fwd(logits, target, mask, P, work, loss):
    B, Cc, H, W = logits.shape
    assert logits.is_contiguous() and target.is_contiguous() and target.dtype == torch.int64 and target.shape == (B, H, W)
    mb = mask.stride(0) if mask is not None else 0
    if mask is not None:
        assert mask.dtype == torch.int64 and mask.stride(1) == 1
    check(_L().mmf_masked_ce_fwd(_p(logits), int(logits.dtype == f32), _p(target), _p(mask), mb, B, Cc, H, W, P, _p(work),
                                 _p(loss), _stream()), "mmf_masked_ce_fwd")


def masked_ce_bwd(logits, target, mask, P, work, dloss, dlogits):
    B, Cc, H, W = logits.shape
    mb = mask.stride(0) if mask is not None else 0
    check(_L().mmf_masked_ce_bwd(_p(logits), int(logits.dtype == f32), _p(target), _p(mask), mb, B, Cc, H, W, P, _p(work),
                                 _p(dloss), _p(dlogits), _stream()), "mmf_masked_ce_bwd")


def cast_bf16(src, dst=None, *, rows_pad=None, cols_pad=None, scale=1.0):
    """f32 [rows, cols] -> bf16 [rows_pad, cols_pad] (zero padded)."""
    assert src.dtype == f32 and src.dim() == 2
    rows, cols = src.shape
    rows_pad = rows_pad or rows
    cols_pad = cols_pad or cols
    if dst is None:
        dst = torch.empty(rows_pad, cols_pad, dtype=bf16, device=src.device)
    check(_L().mmf_cast_f32_bf16(_p(src), rows, cols, _ld(src), _p(dst), rows_pad, cols_pad, _ld(dst), scale, _stream()),
          "mmf_cast_f32_bf16")
    return dst


def geglu_bwd(u, dg, du):
    rows, ipad = dg.shape
    assert u.is_contiguous() and dg.is_contiguous() and du.is_contiguous() and u.shape[1] == 2 * ipad
    check(_L().mmf_geglu_bwd(_p(u), _p(dg), _p(du), rows, ipad, _stream()), "mmf_geglu_bwd")
    return du


def gelu_bwd(pre, dy, dpre):
    assert pre.is_contiguous() and dy.is_contiguous() and dpre.is_contiguous()
    check(_L().mmf_gelu_bwd(_p(pre), _p(dy), _p(dpre), pre.numel(), _stream()), "mmf_gelu_bwd")
    return dpre


def colsum(x, out):
    assert out.dtype == f32
    check(_L().mmf_colsum(_p(x), int(x.dtype == f32), x.shape[0], x.shape[1], _ld(x), _p(out), _stream()), "mmf_colsum")
    return out


def bcast_rows(src, dst, batch, rows, d, dst_batch_stride):
    check(_L().mmf_bcast_rows(_p(src), _p(dst), batch, rows, d, dst_batch_stride, _stream()), "mmf_bcast_rows")


def reduce_batch(src, dst, batch, rows, d, src_batch_stride):
    check(_L().mmf_reduce_batch(_p(src), _p(dst), batch, rows, d, src_batch_stride, _stream()), "mmf_reduce_batch")


def im2col_gather(img, idx, out, P):
    B, Cc, H, W = img.shape
    assert img.is_contiguous() and img.dtype == f32 and idx.dtype == torch.int32 and out.dtype == bf16
    check(_L().mmf_im2col_gather(_p(img), _p(idx), _p(out), B, Cc, H, W, P, idx.numel(), _ld(out), _stream()), "mmf_im2col_gather")
    return out


_RASTER_DTYPES = {torch.uint8: 0, torch.uint16: 1, torch.float32: 2}
RASTER_ZSCORE, RASTER_SAR_DB, RASTER_STANDARDIZE = 0, 1, 2


def raster_prep(src, mode, factor, mean=None, std=None, crop_top=None, crop_left=None, out_hw=None, out=None):
    """raw raster batch [B, C, Hs, Ws] (uint8 / uint16 / float32) -> normalised fp32 [B, C, Ho, Wo]; see mmf_raster_prep"""
    assert src.dim() == 4 and src.is_contiguous() and src.dtype in _RASTER_DTYPES, "raster_prep: contiguous uint8/uint16/float32 [B,C,H,W]"
    if src.data_ptr() % (2 * src.element_size()):
        src = src.clone()          # a view at an odd element offset: the factor-2 path reads pixel pairs with one load
    B, Cc, Hs, Ws = src.shape
    Ho, Wo = out_hw if out_hw is not None else (Hs // factor, Ws // factor)
    if out is None:
        out = torch.empty(B, Cc, Ho, Wo, dtype=f32, device=src.device)
    assert out.is_contiguous() and out.dtype == f32 and tuple(out.shape) == (B, Cc, Ho, Wo)
    for t in (crop_top, crop_left):
        assert t is None or (t.dtype == torch.int32 and t.numel() == B and t.is_contiguous())
    dbl = C.c_double * max(Cc, 1)
    m = dbl(*[float(v) for v in mean]) if mean is not None else None
    sd = dbl(*[float(v) for v in std]) if std is not None else None
    check(_L().mmf_raster_prep(_p(src), _RASTER_DTYPES[src.dtype], B, Cc, Hs, Ws, factor, mode, m, sd, _p(crop_top), _p(crop_left),
                               Ho, Wo, _p(out), _stream()), "mmf_raster_prep")
    return out


def trunc_standardize(x, lo=0.1, hi=0.9):
    """x fp32 [B, ...]: per sample (x - mean) / sqrt(var + 1e-6) with the mean / unbiased variance of the sorted values
    between the `lo` and `hi` quantile positions (pretrain_mmae.py:452-459)"""
    assert x.dtype == f32 and x.is_contiguous() and x.dim() >= 2
    B = x.shape[0]
    n = x.numel() // B
    out = torch.empty_like(x)
    check(_L().mmf_trunc_standardize(_p(x), _p(out), B, n, int(lo * n), int(hi * n), _stream()), "mmf_trunc_standardize")
    return out


def onehot_im2col(cls, idx, out, P, num_classes):
    """cls [B, H, W] int64 class map -> out (zeroed bf16 [B*n, num_classes*P*P]) one-hot rows of the visible patches"""
    B, H, W = cls.shape
    assert cls.is_contiguous() and cls.dtype == torch.int64 and idx.dtype == torch.int32 and out.dtype == bf16
    check(_L().mmf_onehot_im2col(_p(cls), _p(idx), _p(out), B, H, W, P, idx.numel(), num_classes, _ld(out), _stream()),
          "mmf_onehot_im2col")
    return out


def unpatchify(tokens, image, Cc, H, W, P, inverse=False):
    B = image.shape[0]
    assert tokens.is_contiguous() and image.is_contiguous() and tokens.dtype == bf16 and image.dtype == bf16
    check(_L().mmf_unpatchify_bf16(_p(tokens), _p(image), B, Cc, H, W, P, int(inverse), _stream()), "mmf_unpatchify_bf16")


def gather_rows(src, dst, *, batch, n, d, src_batch_rows, row_off=0, idx=None):
    check(_L().mmf_gather_rows(_p(src), int(src.dtype == f32), src.stride(0), src_batch_rows, row_off, _p(idx), _p(dst),
                               int(dst.dtype == f32), dst.stride(0), batch, n, d, _stream()), "mmf_gather_rows")
    return dst


def add_inplace(y, x):
    assert y.dtype == f32 and x.dtype == f32 and y.is_contiguous() and x.is_contiguous() and y.numel() == x.numel()
    check(_L().mmf_add_inplace_f32(_p(y), _p(x), y.numel(), _stream()), "mmf_add_inplace_f32")
    return y


def add_bf16(x, d, out=None):
    """out (f32) = x (f32) + d (bf16)"""
    assert x.dtype == f32 and d.dtype == bf16 and x.is_contiguous() and d.is_contiguous() and x.numel() == d.numel()
    if out is None:
        out = torch.empty_like(x)
    check(_L().mmf_add_bf16_f32(_p(out), _p(x), _p(d), x.numel(), _stream()), "mmf_add_bf16_f32")
    return out


def mask_build(noise1, noise2, share, sizes, nenc, n_fusion, want_slotmap):
    """Everything of generate_random_masks / token selection after the random draws (see include/mmf_b200.h).
    Returns (mask [n] int64, ids_restore [n] int64, ids_keep [nenc] int64, idx [n] int32, counts [T] int32,
    seg [T+2] int32, slotmap [T, n_fusion] int32 or None, tok [nenc] int32), all on the device, no host sync."""
    dev = noise1.device
    T, n = len(sizes), int(sum(sizes))
    assert noise1.dtype == f32 and noise2.dtype == f32 and share.dtype == f32
    assert noise1.numel() == n and noise2.numel() == n and share.numel() == T
    mask = torch.empty(n, dtype=torch.int64, device=dev)
    ids_restore = torch.empty(n, dtype=torch.int64, device=dev)
    ids_keep = torch.empty(nenc, dtype=torch.int64, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.empty(T, dtype=torch.int32, device=dev)
    seg = torch.empty(T + 2, dtype=torch.int32, device=dev)
    slotmap = torch.empty(T, n_fusion, dtype=torch.int32, device=dev) if want_slotmap else None
    tok = torch.empty(nenc, dtype=torch.int32, device=dev)
    csizes = (C.c_int32 * T)(*[int(v) for v in sizes])
    check(_L().mmf_mask_build(_p(noise1), _p(noise2), _p(share), T, csizes, nenc, n_fusion, _p(mask), _p(ids_restore),
                              _p(ids_keep), _p(idx), _p(counts), _p(seg), _p(slotmap), _p(tok), _stream()), "mmf_mask_build")
    return mask, ids_restore, ids_keep, idx, counts, seg, slotmap, tok


def mask_explicit(given, sizes, nenc, n_fusion, want_slotmap):
    """the same bookkeeping for a caller-provided mask row `given` [sum sizes] int64 (0 = visible): returns (ids_restore,
    ids_keep, idx, counts, seg, slotmap, tok, err [1] int32), all on the device, no host sync"""
    dev = given.device
    T, n = len(sizes), int(sum(sizes))
    assert given.dtype == torch.int64 and given.is_contiguous() and given.numel() == n
    ids_restore = torch.empty(n, dtype=torch.int64, device=dev)
    ids_keep = torch.zeros(nenc, dtype=torch.int64, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    counts = torch.empty(T, dtype=torch.int32, device=dev)
    seg = torch.empty(T + 2, dtype=torch.int32, device=dev)
    slotmap = torch.empty(T, n_fusion, dtype=torch.int32, device=dev) if want_slotmap else None
    tok = torch.zeros(nenc, dtype=torch.int32, device=dev)
    err = torch.empty(1, dtype=torch.int32, device=dev)
    csizes = (C.c_int32 * T)(*[int(v) for v in sizes])
    check(_L().mmf_mask_explicit(_p(given), T, csizes, nenc, n_fusion, _p(ids_restore), _p(ids_keep), _p(idx), _p(counts), _p(seg),
                                 _p(slotmap), _p(tok), _p(err), _stream()), "mmf_mask_explicit")
    return ids_restore, ids_keep, idx, counts, seg, slotmap, tok, err


def im2col_tokens(imgs, tok, out, P, col_off, tok_off, ind_col):
    """token-table im2col over all modalities (see mmf_im2col_tokens): imgs = list of [B, C_m, H, W] f32, tok [nenc] int32"""
    M = len(imgs)
    B, _, H, W = imgs[0].shape
    for im in imgs:
        assert im.is_contiguous() and im.dtype == f32 and im.shape[0] == B and tuple(im.shape[2:]) == (H, W)
    assert tok.dtype == torch.int32 and out.dtype == bf16 and out.shape[0] == B * tok.numel()
    ptrs = (C.c_void_p * M)(*[_p(im) for im in imgs])
    chans = (C.c_int32 * M)(*[int(im.shape[1]) for im in imgs])
    coff = (C.c_int32 * M)(*[int(v) for v in col_off])
    toff = (C.c_int32 * (M + 1))(*[int(v) for v in tok_off])
    check(_L().mmf_im2col_tokens(ptrs, chans, coff, toff, M, _p(tok), tok.numel(), _p(out), _ld(out), int(ind_col), B, H, W, P, _stream()),
          "mmf_im2col_tokens")
    return out


def dino_loss(student, teacher, student_temp, teacher_temp):
    """-> (row_loss [B] f32, dstudent [B, D] f32 for an upstream gradient of 1)"""
    assert student.dim() == 2 and student.shape == teacher.shape and student.dtype == teacher.dtype
    assert student.dtype in (bf16, f32) and student.stride(1) == 1 and teacher.stride(1) == 1
    B, D = student.shape
    row_loss = torch.empty(B, dtype=f32, device=student.device)
    dstudent = torch.empty(B, D, dtype=f32, device=student.device)
    check(_L().mmf_dino_loss(_p(student), student.stride(0), _p(teacher), teacher.stride(0), int(student.dtype == f32), B, D,
                             student_temp, teacher_temp, _p(row_loss), _p(dstudent), _stream()), "mmf_dino_loss")
    return row_loss, dstudent


def hardneg_loss(out1, out2, tau_plus, beta, temperature, easy=False):
    """-> (loss [1] f32, dout1, dout2 [B, D] f32 for an upstream gradient of 1); out1 / out2: [B, D] f32 rows"""
    assert out1.dim() == 2 and out1.shape == out2.shape and out1.dtype == f32 and out2.dtype == f32
    assert out1.stride(1) == 1 and out2.stride(1) == 1
    B, D = out1.shape
    L = _L()
    work = torch.empty(int(L.mmf_hardneg_workspace_floats(B, D)), dtype=f32, device=out1.device)
    loss = torch.empty(1, dtype=f32, device=out1.device)
    d1 = torch.empty(B, D, dtype=f32, device=out1.device)
    d2 = torch.empty(B, D, dtype=f32, device=out1.device)
    check(L.mmf_hardneg_loss(_p(out1), out1.stride(0), _p(out2), out2.stride(0), B, D, float(tau_plus), float(beta), float(temperature),
                             int(easy), _p(work), _p(loss), _p(d1), _p(d2), _stream()), "mmf_hardneg_loss")
    return loss, d1, d2
