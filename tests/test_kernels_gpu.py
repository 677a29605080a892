"""GPU: every C-ABI kernel against a plain PyTorch fp32 reference of the same op (through the C ABI)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


def K():
    from incomplete_multimodal_fusion_b200 import kernels
    return kernels


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def rnd(*shape, dtype=f32, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K_", [(128, 128, 64), (256, 256, 128), (300, 200, 136), (1000, 768, 512), (77, 1536, 768),
                                    (513, 96, 64)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_layouts(M, N, K_, a_mn, b_mn):
    # MN-major operands need 16-byte row pitch on the contiguous (M or N) axis
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("pitch not 16B aligned for this layout")
    A = rnd(M, K_, dtype=bf16, seed=1)
    B = rnd(N, K_, dtype=bf16, seed=2)
    ref = A.float() @ B.float().t()
    a = A.t().contiguous() if a_mn else A
    b = B.t().contiguous() if b_mn else B
    for out_dtype in (bf16, f32):
        out = torch.empty(M, N, dtype=out_dtype, device="cuda")
        K().gemm(a, b, out, a_mn=a_mn, b_mn=b_mn)
        torch.cuda.synchronize()
        assert rel(out, ref) < (6e-3 if out_dtype == bf16 else 1e-5), (out_dtype, rel(out, ref))


@pytest.mark.parametrize("block_n", [128, 256])
def test_gemm_epilogues(block_n):
    M, N, K_ = 700, 512, 256
    A = rnd(M, K_, dtype=bf16, seed=3)
    W = rnd(N, K_, dtype=bf16, seed=4, scale=0.1)
    bias = rnd(N, seed=5)
    res = rnd(M, N, seed=6)
    acc = A.float() @ W.float().t()
    # bias + gelu, bf16 out, pre-activation kept
    out = torch.empty(M, N, dtype=bf16, device="cuda")
    pre = torch.empty(M, N, dtype=bf16, device="cuda")
    K().gemm(A, W, out, bias=bias, act=1, out2=pre, block_n=block_n)
    assert rel(out, F.gelu(acc + bias)) < 6e-3
    assert rel(pre, acc + bias) < 6e-3
    # bias + residual, f32 out + bf16 copy, alpha
    out = torch.empty(M, N, dtype=f32, device="cuda")
    cp = torch.empty(M, N, dtype=bf16, device="cuda")
    K().gemm(A, W, out, bias=bias, residual=res, alpha=0.5, out2=cp, block_n=block_n)
    ref = 0.5 * acc + bias + res
    assert rel(out, ref) < 1e-5
    assert rel(cp, ref) < 6e-3
    # two-source residual
    r1, r2 = res[:300].contiguous(), res[300:].contiguous()
    K().gemm(A, W, out, residual=r1, residual2=r2, res_split=300, block_n=block_n)
    assert rel(out, acc + res) < 1e-5
    # residual through a row map with a period (pos-emb gather) + batched output rows
    period, nb = 70, 10
    table = rnd(100, N, seed=7)
    rmap = torch.randperm(100, device="cuda")[:period].int()
    big = torch.zeros(nb * 90, N, dtype=f32, device="cuda")
    K().gemm(A, W, big[5:], residual=table, res_row_map=rmap, res_period=period, out_period=period, out_batch_rows=90,
             block_n=block_n)
    ref = (acc.view(nb, period, N) + table[rmap.long()][None]).reshape(M, N)
    got = big.view(nb, 90, N)[:, 5:5 + period].reshape(M, N)
    assert rel(got, ref) < 1e-5
    assert float(big.view(nb, 90, N)[:, :5].abs().max()) == 0.0


def test_gemm_split_k_wgrad():
    # dW[N_out, K_in] = dY^T X with both operands token-major (MN-major), huge K = tokens
    T, NO, KI = 5000, 384, 256
    dY = rnd(T, NO, dtype=bf16, seed=8)
    X = rnd(T, KI, dtype=bf16, seed=9)
    ref = dY.float().t() @ X.float()
    for split in (1, 4, 7):
        out = torch.zeros(NO, KI, dtype=f32, device="cuda")
        K().gemm(dY, X, out, a_mn=True, b_mn=True, split_k=split)
        assert rel(out, ref) < 1e-5, split


@pytest.mark.parametrize("I", [512, 344, 2048])
def test_gemm_geglu(I):
    M, D = 333, 256
    X = rnd(M, D, dtype=bf16, seed=10)
    W = rnd(2 * I, D, dtype=bf16, seed=11, scale=0.1)
    u = X.float() @ W.float().t()
    ref = F.gelu(u[:, I:]) * u[:, :I]
    out = torch.empty(M, I, dtype=bf16, device="cuda")
    pre = torch.empty(M, 2 * I, dtype=bf16, device="cuda")
    K().gemm(X, W, out, act=2, out2=pre)
    assert rel(out, ref) < 6e-3
    assert rel(pre, u) < 6e-3


@pytest.mark.parametrize("M,I", [(333, 512), (4096 + 77, 2048), (1000, 2752), (256, 64), (40000, 1024)])
def test_gemm_geglu_backward_fused(M, I):
    # du = [dg * gelu(gate) | dg * value * gelu'(gate)] with dg = dY . W2 formed in TMEM only (act=3), against
    # autograd of the reference expression (zorro_utils.py:115-118) in fp32 and against the unfused kernels
    D = 256
    dY = rnd(M, D, dtype=bf16, seed=20, scale=0.5)
    W2 = rnd(D, I, dtype=bf16, seed=21, scale=0.1)          # stored [K = D, N = I]: the dgrad layout (b_mn)
    u = rnd(M, 2 * I, dtype=bf16, seed=22)
    uf = u.float().requires_grad_(True)
    dg = dY.float() @ W2.float()
    (F.gelu(uf[:, I:]) * uf[:, :I]).backward(dg)
    du = torch.full((M, 2 * I), float("nan"), dtype=bf16, device="cuda")
    K().gemm(dY, W2, du, b_mn=True, act=3, out2=u)
    assert torch.isfinite(du.float()).all()
    assert rel(du, uf.grad) < 6e-3
    dg_b = torch.empty(M, I, dtype=bf16, device="cuda")
    K().gemm(dY, W2, dg_b, b_mn=True)
    du2 = K().geglu_bwd(u, dg_b, torch.empty_like(u))
    assert rel(du, du2) < 8e-3


def test_gemm_large_persistent():
    # more tiles than SMs: exercises the persistent loop, both accumulator stages and ring wrap-around
    M, N, K_ = 4096 + 64, 2048, 1024
    A = rnd(M, K_, dtype=bf16, seed=12)
    B = rnd(N, K_, dtype=bf16, seed=13)
    out = torch.empty(M, N, dtype=bf16, device="cuda")
    K().gemm(A, B, out)
    assert rel(out, A.float() @ B.float().t()) < 6e-3


# ------------------------------------------------------------------------------------------------
# LayerNorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [128, 192, 256, 768, 1024])
@pytest.mark.parametrize("double", [False, True])
def test_layernorm_fwd_bwd(D, double):
    rows = 517
    x = rnd(rows, D, seed=1, scale=2.0).requires_grad_(True)
    g1 = (1 + 0.1 * rnd(D, seed=2)).requires_grad_(True)
    b1 = None if double else (0.1 * rnd(D, seed=3)).requires_grad_(True)
    g2 = (1 + 0.1 * rnd(D, seed=4)).requires_grad_(True) if double else None
    eps1, eps2 = (1e-5, 1e-5) if double else (1e-6, 0.0)
    y_ref = F.layer_norm(x, (D,), g1, b1, eps1)
    if double:
        y_ref = F.layer_norm(y_ref, (D,), g2, None, eps2)
    dy = rnd(rows, D, seed=5)
    dres = rnd(rows, D, seed=6)
    y_ref.backward(dy)
    for ydt in (f32, bf16):
        y = torch.empty(rows, D, dtype=ydt, device="cuda")
        stats = torch.empty(rows, 4, device="cuda")
        K().layernorm_fwd(x.detach(), g1.detach(), y, b1=None if b1 is None else b1.detach(), eps1=eps1,
                          g2=None if g2 is None else g2.detach(), eps2=eps2, stats=stats)
        assert rel(y, y_ref) < (1e-5 if ydt == f32 else 5e-3)
    dx = torch.empty(rows, D, device="cuda")
    dxb = torch.empty(rows, D, dtype=bf16, device="cuda")
    dg1 = torch.zeros(D, device="cuda"); db1 = torch.zeros(D, device="cuda"); dg2 = torch.zeros(D, device="cuda")
    K().layernorm_bwd(dy, x.detach(), g1.detach(), stats, dx, dg1, b1=None if b1 is None else b1.detach(),
                      g2=None if g2 is None else g2.detach(), dres=dres, dx_bf16=dxb, db1=None if b1 is None else db1,
                      dg2=dg2 if double else None)
    assert rel(dx, x.grad + dres) < 2e-5
    assert rel(dxb, x.grad + dres) < 5e-3
    assert rel(dg1, g1.grad) < 2e-5
    if b1 is not None:
        assert rel(db1, b1.grad) < 2e-5
    if double:
        assert rel(dg2, g2.grad) < 2e-5


@pytest.mark.parametrize("D", [256, 512, 768, 1024])
@pytest.mark.parametrize("double", [False, True])
@pytest.mark.parametrize("rows", [300, 9001])
def test_layernorm_bwd_ring(D, double, rows, monkeypatch):
    """the bulk-copy ring kernel (bf16 upstream gradient, D a multiple of 128): against torch autograd, and bit-identical dx to
    the register-prefetch kernel it replaces (MMF_LN_BWD_RING=0); 9001 rows = up to five rows per warp, so both stages of the
    ring wrap and the barrier phase flips; with and without the residual-branch gradient, with the input in two buffers"""
    x = rnd(rows, D, seed=1, scale=2.0).requires_grad_(True)
    g1 = (1 + 0.1 * rnd(D, seed=2)).requires_grad_(True)
    b1 = None if double else (0.1 * rnd(D, seed=3)).requires_grad_(True)
    g2 = (1 + 0.1 * rnd(D, seed=4)).requires_grad_(True) if double else None
    eps1, eps2 = (1e-5, 1e-5) if double else (1e-6, 0.0)
    y_ref = F.layer_norm(x, (D,), g1, b1, eps1)
    if double:
        y_ref = F.layer_norm(y_ref, (D,), g2, None, eps2)
    dy = rnd(rows, D, seed=5).to(bf16)
    dres = rnd(rows, D, seed=6)
    y_ref.backward(dy.float())
    y = torch.empty(rows, D, dtype=bf16, device="cuda")
    stats = torch.empty(rows, 4, device="cuda")
    kw = dict(b1=None if b1 is None else b1.detach(), g2=None if g2 is None else g2.detach())
    K().layernorm_fwd(x.detach(), g1.detach(), y, eps1=eps1, eps2=eps2, stats=stats, **kw)
    split = rows // 3
    xa, xb = x.detach()[:split].clone(), x.detach()[split:].clone()
    outs = {}
    for ring in ("1", "0"):
        monkeypatch.setenv("MMF_LN_BWD_RING", ring)
        for case in ("dres", "plain", "two_sources"):
            dx = torch.full((rows, D), float("nan"), device="cuda")
            dxb = torch.full((rows, D), float("nan"), dtype=bf16, device="cuda")
            dg1 = torch.zeros(D, device="cuda"); db1 = torch.zeros(D, device="cuda"); dg2 = torch.zeros(D, device="cuda")
            src = dict(x2=xb, x_split=split, rows=rows) if case == "two_sources" else {}
            K().layernorm_bwd(dy, xa if case == "two_sources" else x.detach(), g1.detach(), stats, dx, dg1,
                              dres=dres if case == "dres" else None, dx_bf16=dxb, db1=None if b1 is None else db1,
                              dg2=dg2 if double else None, **kw, **src)
            want = x.grad + dres if case == "dres" else x.grad
            assert rel(dx, want) < 2e-5 and rel(dxb, want) < 5e-3, (ring, case)
            assert rel(dg1, g1.grad) < 2e-5, (ring, case)
            if b1 is not None:
                assert rel(db1, b1.grad) < 2e-5, (ring, case)
            if double:
                assert rel(dg2, g2.grad) < 2e-5, (ring, case)
            outs[ring, case] = (dx, dxb)
    for case in ("dres", "plain", "two_sources"):
        assert torch.equal(outs["1", case][0], outs["0", case][0]) and torch.equal(outs["1", case][1], outs["0", case][1]), case


def test_layernorm_two_sources():
    D = 256
    xa, xb = rnd(100, D, seed=1), rnd(60, D, seed=2)
    g = 1 + 0.1 * rnd(D, seed=3)
    y = torch.empty(160, D, device="cuda")
    K().layernorm_fwd(xa, g, y, x2=xb, x_split=100, rows=160)
    ref = F.layer_norm(torch.cat([xa, xb]), (D,), g, None, 1e-5)
    assert rel(y, ref) < 1e-5


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def _planar(x, B, n_head):
    """[B, N, C] token-major -> planar rows (all head tokens of the batch, then all tail tokens)"""
    return torch.cat([x[:, :n_head].reshape(-1, x.shape[-1]), x[:, n_head:].reshape(-1, x.shape[-1])], 0).contiguous()


def _unplanar(x, B, N, n_head):
    Cc = x.shape[-1]
    return torch.cat([x[:B * n_head].view(B, n_head, Cc), x[B * n_head:].view(B, N - n_head, Cc)], 1)


@pytest.mark.parametrize("dh,H", [(64, 2), (32, 4)])
@pytest.mark.parametrize("counts,nf", [((40, 70, 18), 64), ((0, 100, 28), 64), ((64, 64, 0), 100), (None, 150)])
def test_attention_fwd_bwd(dh, H, counts, nf):
    B = 3
    if counts is None:
        N, n_head, seg, nseg, allowed = nf, nf, None, 0, None
    else:
        nenc = sum(counts)
        N, n_head = nenc + nf, nenc
        bounds = [0, counts[0], counts[0] + counts[1], nenc, N]
        seg = torch.tensor(bounds, dtype=torch.int32, device="cuda")
        nseg = 4
        types = torch.cat([torch.full((c,), i) for i, c in enumerate(list(counts) + [nf])]).cuda()
        allowed = (types[:, None] == types[None, :]) | (types[:, None] == 3)
    scale = dh ** -0.5
    qkv = rnd(B, N, 3 * H * dh, dtype=bf16, seed=1).requires_grad_(True)
    q, k, v = [t.view(B, N, H, dh).transpose(1, 2).float() for t in qkv.chunk(3, -1)]
    s = (q @ k.transpose(-1, -2)) * scale
    if allowed is not None:
        s = s.masked_fill(~allowed, float("-inf"))
    o_ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * dh)
    do = rnd(B, N, H * dh, dtype=bf16, seed=2)
    o_ref.backward(do.float())

    qkv_p = _planar(qkv.detach(), B, n_head)
    HD = H * dh
    o = torch.empty(B * N, HD, dtype=bf16, device="cuda")
    lse = torch.empty(B, H, N, device="cuda")
    kw = dict(B=B, H=H, Nq=N, Nk=N, dh=dh, scale=scale, n_head_q=n_head, n_head_k=n_head, seg=seg, nseg=nseg)
    K().attn_fwd(qkv_p[:, :HD], qkv_p[:, HD:2 * HD], qkv_p[:, 2 * HD:], o, lse, **kw)
    assert rel(_unplanar(o, B, N, n_head), o_ref) < 8e-3
    dqkv = torch.full_like(qkv_p, float("nan"))
    delta = torch.empty(B, H, N, device="cuda")
    K().attn_bwd(qkv_p[:, :HD], qkv_p[:, HD:2 * HD], qkv_p[:, 2 * HD:], o, lse, _planar(do, B, n_head),
                 dqkv[:, :HD], dqkv[:, HD:2 * HD], dqkv[:, 2 * HD:], delta, **kw)
    got = _unplanar(dqkv, B, N, n_head)
    assert torch.isfinite(got.float()).all()
    assert rel(got, qkv.grad) < 1.5e-2


@pytest.mark.parametrize("counts", [(98, 98, 98), (150, 0, 144), (37, 196, 61), (0, 196, 98), (148, 148, 148), (196, 196, 196),
                                    (1, 196, 97), (294, 0, 0)])
def test_attention_encoder_shapes(counts):
    """the tcgen05 kernels at the encoder's own shapes: H = 8 heads of 64, 196 fusion tokens behind nenc = 294 / 444 / 588
    visible tokens (N = 490 / 640 / 784) in ragged modality segments, including an absent modality, a one-token segment
    and a single-modality sample; per-row errors next to the L2 ratio (a few wrong rows would hide in the latter)"""
    B, H, dh, nf = 2, 8, 64, 196
    nenc = sum(counts)
    N, n_head = nenc + nf, nenc
    seg = torch.tensor([0, counts[0], counts[0] + counts[1], nenc, N], dtype=torch.int32, device="cuda")
    types = torch.cat([torch.full((c,), i) for i, c in enumerate(list(counts) + [nf])]).cuda()
    allowed = (types[:, None] == types[None, :]) | (types[:, None] == 3)
    scale = dh ** -0.5
    qkv = rnd(B, N, 3 * H * dh, dtype=bf16, seed=11).requires_grad_(True)
    q, k, v = [t.view(B, N, H, dh).transpose(1, 2).float() for t in qkv.chunk(3, -1)]
    s = ((q @ k.transpose(-1, -2)) * scale).masked_fill(~allowed, float("-inf"))
    o_ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, H * dh)
    do = rnd(B, N, H * dh, dtype=bf16, seed=12)
    o_ref.backward(do.float())
    qkv_p = _planar(qkv.detach(), B, n_head)
    HD = H * dh
    o = torch.full((B * N, HD), float("nan"), dtype=bf16, device="cuda")
    lse = torch.empty(B, H, N, device="cuda")
    kw = dict(B=B, H=H, Nq=N, Nk=N, dh=dh, scale=scale, n_head_q=n_head, n_head_k=n_head, seg=seg, nseg=4)
    K().attn_fwd(qkv_p[:, :HD], qkv_p[:, HD:2 * HD], qkv_p[:, 2 * HD:], o, lse, **kw)
    got_o = _unplanar(o, B, N, n_head).float()
    assert torch.isfinite(got_o).all()
    assert rel(got_o, o_ref) < 8e-3
    lse_ref = torch.logsumexp(s, -1)                                     # [B, H, N] in token order
    assert (lse - lse_ref).abs().max() < 2e-2
    row = lambda a, b: ((a - b).norm(dim=-1) / b.norm(dim=-1).mean()).max()
    assert row(got_o, o_ref.detach()) < 4e-2
    dqkv = torch.full_like(qkv_p, float("nan"))
    delta = torch.empty(B, H, N, device="cuda")
    K().attn_bwd(qkv_p[:, :HD], qkv_p[:, HD:2 * HD], qkv_p[:, 2 * HD:], o, lse, _planar(do, B, n_head),
                 dqkv[:, :HD], dqkv[:, HD:2 * HD], dqkv[:, 2 * HD:], delta, **kw)
    got = _unplanar(dqkv, B, N, n_head).float()
    assert torch.isfinite(got).all()
    assert rel(got, qkv.grad) < 1.5e-2
    for part, name in zip(range(3), "qkv"):
        a, b = got[..., part * HD:(part + 1) * HD], qkv.grad[..., part * HD:(part + 1) * HD]
        assert rel(a, b) < 1.5e-2, name
        assert row(a, b) < 8e-2, name


def test_slot_attention_fwd_bwd():
    B, Fn, H, S = 3, 16, 2, 4
    counts = (9, 0, 5)
    nenc = sum(counts)
    HD = H * 64
    g = torch.Generator().manual_seed(0)
    idx = [torch.randperm(Fn, generator=g)[:c].sort().values for c in counts]
    slotmap = torch.full((3, Fn), -1, dtype=torch.int32)
    for m, ix in enumerate(idx):
        slotmap[m, ix] = torch.arange(len(ix), dtype=torch.int32)
    seg = torch.tensor([0, counts[0], counts[0] + counts[1], nenc], dtype=torch.int32)
    q = rnd(B * Fn, HD, dtype=bf16, seed=1).requires_grad_(True)
    kv_tok = rnd(B * nenc + B * Fn, 2 * HD, dtype=bf16, seed=2).requires_grad_(True)
    kv_me = rnd(Fn, 2 * HD, dtype=bf16, seed=3).requires_grad_(True)
    # reference: gather the S slot rows per (b, p)
    rows = []
    for s_ in range(3):
        r = slotmap[s_].long()
        tok = torch.stack([kv_tok.float()[b * nenc + seg[s_] + r.clamp_min(0).cuda()] for b in range(B)])   # [B, F, 2HD]
        me = kv_me.float()[None].expand(B, -1, -1)
        rows.append(torch.where((r >= 0).cuda()[None, :, None], tok, me))
    rows.append(kv_tok.float()[B * nenc:].view(B, Fn, 2 * HD))
    kvs = torch.stack(rows, 2)                                      # [B, F, S, 2HD]
    kk = kvs[..., :HD].reshape(B, Fn, S, H, 64)
    vv = kvs[..., HD:].reshape(B, Fn, S, H, 64)
    qq = q.float().view(B, Fn, H, 64)
    sc = torch.einsum("bfhd,bfshd->bfhs", qq, kk) * 0.125
    pr = sc.softmax(-1)
    ref = torch.einsum("bfhs,bfshd->bfhd", pr, vv).reshape(B * Fn, HD)
    dout = rnd(B * Fn, HD, dtype=bf16, seed=4)
    ref.backward(dout.float())

    out = torch.empty(B * Fn, HD, dtype=bf16, device="cuda")
    probs = torch.empty(B * Fn, H, S, device="cuda")
    kw = dict(B=B, F=Fn, H=H, S=S, n_head=nenc, scale=0.125)
    sm, sg = slotmap.cuda(), seg.cuda()
    K().slot_attn_fwd(q.detach(), kv_tok.detach(), kv_me.detach(), sm, sg, out, probs, **kw)
    assert rel(out, ref) < 6e-3
    assert rel(probs.view(B, Fn, H, S), pr) < 1e-4
    out2 = torch.empty_like(out)                              # without probs: the warp-per-position kernel
    K().slot_attn_fwd(q.detach(), kv_tok.detach(), kv_me.detach(), sm, sg, out2, None, **kw)
    assert rel(out2, ref) < 6e-3
    dq = torch.empty_like(out)
    dkv = torch.full_like(kv_tok.detach(), float("nan"))
    dme = torch.full((Fn, 2 * HD), float("nan"), device="cuda")   # the batch-reduction kernel writes every element
    K().slot_attn_bwd(q.detach(), kv_tok.detach(), kv_me.detach(), sm, sg, dout, dq, dkv, dme, **kw)
    assert torch.isfinite(dkv.float()).all()
    assert rel(dq, q.grad) < 1e-2
    assert rel(dkv, kv_tok.grad) < 1e-2
    assert rel(dme, kv_me.grad) < 1e-2


def test_pool_attention_fwd_bwd_uniform_rows():
    B, H, R = 3, 2, 5
    counts, nf = (30, 0, 21), 40
    nenc = sum(counts); N = nenc + nf
    HD = H * 64
    types = torch.cat([torch.full((c,), i) for i, c in enumerate(list(counts) + [nf])]).cuda()
    rtypes = torch.tensor([0, 1, 2, 3, 1]).cuda()
    mask = (rtypes[:, None] == types[None, :]) | (rtypes[:, None] == 3)       # row 1 and 4: no allowed key
    mode = torch.tensor([0, 0, 0, 0, 1], dtype=torch.int32, device="cuda")    # row 1 -> uniform, row 4 -> zero
    q = rnd(R, HD, dtype=bf16, seed=1).requires_grad_(True)
    kv = rnd(B, N, 2 * HD, dtype=bf16, seed=2).requires_grad_(True)
    qq = q.float().view(R, H, 64).transpose(0, 1)[None]                        # [1, H, R, 64]
    kk = kv.float()[..., :HD].view(B, N, H, 64).transpose(1, 2)
    vv = kv.float()[..., HD:].view(B, N, H, 64).transpose(1, 2)
    sim = (qq * 0.125) @ kk.transpose(-1, -2)
    sim = sim.masked_fill(~mask, -torch.finfo(sim.dtype).max)
    pr = sim.softmax(-1)
    zero_rows = torch.tensor([False, False, False, False, True]).cuda()
    pr = torch.where(zero_rows[None, None, :, None], torch.zeros_like(pr), pr)
    ref = (pr @ vv).transpose(1, 2).reshape(B, R, HD)
    dout = rnd(B, R, HD, dtype=bf16, seed=3)
    ref.backward(dout.float())

    kvp = _planar(kv.detach(), B, nenc)
    out = torch.empty(B, R, HD, dtype=bf16, device="cuda")
    stat = torch.empty(B * R * H * 3, device="cuda")
    kw = dict(B=B, R=R, H=H, N=N, n_head=nenc, scale=0.125, q_batched=False)
    m8 = mask.to(torch.uint8).contiguous()
    K().pool_attn_fwd(q.detach(), kvp, m8, mode, out, stat, **kw)
    assert rel(out, ref) < 6e-3
    dq = torch.zeros(R, HD, device="cuda")
    dkv = torch.full_like(kvp, float("nan"))
    K().pool_attn_bwd(q.detach(), kvp, m8, mode, out, stat, dout, dq, dkv, **kw)
    assert torch.isfinite(dkv.float()).all()
    assert rel(dq, q.grad) < 1e-2
    assert rel(_unplanar(dkv, B, N, nenc), kv.grad) < 1e-2


# ------------------------------------------------------------------------------------------------
# losses and data movement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("pdt", [bf16, f32])
@pytest.mark.parametrize("P", [8, 4, 16])       # 8 / 16: the vectorised kernels (8 pixels per thread); 4: the scalar ones
def test_masked_loss(kind, pdt, P):
    import oracle
    B, Cc, H, W = 5, 3, 32, 32
    pred = rnd(B, Cc, H, W, dtype=pdt, seed=1).requires_grad_(True)
    tgt = rnd(B, Cc, H, W, seed=2)
    mask = (torch.rand(B, (H // P) * (W // P), device="cuda") > 0.5).long()
    mask[1] = 0                                      # nanmean path
    fn = oracle.masked_mse_loss if kind == 0 else oracle.masked_l1_loss
    for m in (mask, None, torch.zeros_like(mask)):
        pred.grad = None
        ref = fn(pred.float(), tgt, m, P)
        work = torch.empty(2 * B + 2, device="cuda")
        loss = torch.empty(1, device="cuda")
        K().masked_loss_fwd(pred.detach(), tgt, m, P, kind, work, loss)
        assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
        if ref.requires_grad:
            (ref * 1.7).backward()
            dpred = torch.empty_like(pred.detach())
            K().masked_loss_bwd(pred.detach(), tgt, m, P, kind, work, torch.tensor([1.7], device="cuda"), dpred)
            # a sample whose mask is all zero is 0/0 in the reference: nanmean drops it from the loss but
            # autograd still sends NaN (0 * inf) to its pixels; the kernel sends 0.  Compare the other samples.
            keep = torch.ones(B, dtype=torch.bool, device="cuda") if m is None else m.sum(1) > 0
            if (~keep).any():
                assert float(dpred[~keep].float().abs().max()) == 0.0
            assert rel(dpred[keep], pred.grad[keep]) < (1e-5 if pdt == f32 else 6e-3)


@pytest.mark.parametrize("pdt", [bf16, f32])
@pytest.mark.parametrize("C,P", [(9, 16), (5, 8), (33, 8)])
def test_masked_cross_entropy(pdt, C, P, golden_dir):
    """MaskedCrossEntropyLoss (criterion.py:24-58) fused: value and gradient against the oracle restatement (itself pinned by
    the reference's golden in test_losses_match_reference), masked / unmasked / all-zero-mask, through the drop-in class"""
    import oracle
    from incomplete_multimodal_fusion_b200.multimae.criterion import MaskedCrossEntropyLoss
    B, H, W = 4, 32, 32
    logits = rnd(B, C, H, W, dtype=pdt, seed=3, scale=2.0).requires_grad_(True)
    target = torch.randint(0, C, (B, H, W), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    mask = (torch.rand(B, (H // P) * (W // P), device="cuda") > 0.4).long()
    mask[2] = 0                                      # nanmean path
    crit = MaskedCrossEntropyLoss(patch_size=P)
    for m in (mask, None, torch.zeros_like(mask)):
        lref = logits.detach().float().requires_grad_(True)
        ref = oracle.masked_ce_loss(lref, target, m, P)
        logits.grad = None
        out = crit(logits, target, mask=m)
        assert abs(float(out) - float(ref)) <= 2e-5 * max(1.0, abs(float(ref)))
        if ref.requires_grad:
            (ref * 1.3).backward()
            (out * 1.3).backward()
            keep = torch.ones(B, dtype=torch.bool, device="cuda") if m is None else m.sum(1) > 0
            if (~keep).any():      # 0/0 sample: the reference's autograd sends NaN, the kernel 0 (as for the MSE / L1 losses)
                assert float(logits.grad[~keep].float().abs().max()) == 0.0
            assert rel(logits.grad[keep], lref.grad[keep]) < (6e-3 if pdt == bf16 else 1e-5)
    with pytest.raises(NotImplementedError):
        MaskedCrossEntropyLoss(patch_size=P, label_smoothing=0.1)
    # and the reference's golden value on the same inputs
    import os
    fx = torch.load(os.path.join(golden_dir, "losses.pt"), weights_only=False)
    g = torch.Generator().manual_seed(fx["seed"])
    _ = (torch.randn(3, 2, 32, 32, generator=g), torch.randn(3, 2, 32, 32, generator=g))
    gmask = (torch.rand(3, 16, generator=g) > 0.5).long()
    gmask[2] = 0
    _ = (torch.randn(6, 48, generator=g), torch.randn(6, 48, generator=g))
    glogits = torch.randn(3, 5, 32, 32, generator=g)
    gcls = torch.randint(0, 5, (3, 32, 32), generator=g)
    got = MaskedCrossEntropyLoss(patch_size=8)(glogits.cuda(), gcls.cuda(), mask=gmask.cuda())
    assert abs(float(got) - float(fx["ce"])) < 2e-5 * float(fx["ce"])


def test_elementwise():
    import oracle
    k = K()
    # cast with padding
    w = rnd(341, 128, seed=1)
    o = k.cast_bf16(w, rows_pad=384, cols_pad=128)
    assert torch.equal(o[:341], w.to(bf16)) and float(o[341:].abs().max()) == 0
    assert torch.equal(k.cast_bf16(w), w.to(bf16))
    # geglu bwd
    rows, I = 77, 64
    u = rnd(rows, 2 * I, dtype=bf16, seed=2).float().requires_grad_(True)
    dg = rnd(rows, I, dtype=bf16, seed=3)
    (F.gelu(u[:, I:]) * u[:, :I]).backward(dg.float())
    du = torch.empty(rows, 2 * I, dtype=bf16, device="cuda")
    k.geglu_bwd(u.detach().to(bf16), dg, du)
    assert rel(du, u.grad) < 6e-3
    # gelu bwd
    pre = rnd(50, 64, dtype=bf16, seed=4).float().requires_grad_(True)
    dy = rnd(50, 64, dtype=bf16, seed=5)
    F.gelu(pre).backward(dy.float())
    dp = torch.empty(50, 64, dtype=bf16, device="cuda")
    k.gelu_bwd(pre.detach().to(bf16), dy, dp)
    assert rel(dp, pre.grad) < 6e-3
    # colsum
    x = rnd(999, 200, dtype=bf16, seed=6)
    out = torch.zeros(200, device="cuda")
    k.colsum(x, out)
    assert rel(out, x.float().sum(0)) < 1e-5
    for rows, cols in ((5001, 256), (777, 768), (31, 1024), (50176, 256)):   # the vectorised kernel (cols % 256 == 0)
        x = rnd(rows, cols, dtype=bf16, seed=8)
        out = torch.zeros(cols, device="cuda")
        k.colsum(x, out)
        assert rel(out, x.float().sum(0)) < 1e-5, (rows, cols)
    # bcast / reduce
    src = rnd(10, 64, seed=7)
    dst = torch.zeros(4, 25, 64, device="cuda")
    k.bcast_rows(src, dst[:, 15:], 4, 10, 64, 25 * 64)
    assert torch.equal(dst[:, 15:], src[None].expand(4, -1, -1)) and float(dst[:, :15].abs().max()) == 0
    red = torch.empty(10, 64, device="cuda")
    k.reduce_batch(dst[:, 15:], red, 4, 10, 64, 25 * 64)
    assert rel(red, 4 * src) < 1e-6
    # im2col gather == conv patch projection operand
    B, Cc, H, W, P = 2, 3, 32, 32, 8
    img = rnd(B, Cc, H, W, seed=8)
    idx = torch.tensor([0, 3, 7, 12, 15], dtype=torch.int32, device="cuda")
    A = torch.empty(B * 5, Cc * P * P, dtype=bf16, device="cuda")
    k.im2col_gather(img, idx, A, P)
    ref = F.unfold(img, P, stride=P).transpose(1, 2)[:, idx.long()].reshape(B * 5, -1)
    assert torch.equal(A, ref.to(bf16))
    # unpatchify round trip vs oracle
    cfg = oracle.OracleConfig(patch=P, image_size=H)
    tok = rnd(B, 16, Cc * P * P, dtype=bf16, seed=9)
    im = torch.empty(B, Cc, H, W, dtype=bf16, device="cuda")
    k.unpatchify(tok, im, Cc, H, W, P)
    assert torch.equal(im, oracle.functional._unpatchify(tok, cfg, Cc, H, W))
    back = torch.empty_like(tok)
    k.unpatchify(back, im, Cc, H, W, P, inverse=True)
    assert torch.equal(back, tok)
    # gather rows with cast
    src = rnd(3 * 20, 64, seed=10)
    gi = torch.tensor([1, 4, 9], dtype=torch.int32, device="cuda")
    dst = torch.empty(9, 64, dtype=bf16, device="cuda")
    k.gather_rows(src, dst, batch=3, n=3, d=64, src_batch_rows=20, row_off=2, idx=gi)
    assert torch.equal(dst, src.view(3, 20, 64)[:, (gi.long() + 2)].reshape(9, 64).to(bf16))


@pytest.mark.parametrize("sizes,nenc", [((196, 196, 196), 294), ((64, 64), 64), ((256, 256, 256, 256), 1000), ((16, 16, 16), 1),
                                        ((49, 49, 49), 147)])
def test_mask_build_matches_torch_ops(sizes, nenc):
    """the single-CTA mask builder against the reference's op sequence (multimae.py:210-255, 378-382) on the same noise:
    int64 masks / ids bit-exact, index lists = nonzero(), counts, segment table, slot map"""
    T, n = len(sizes), sum(sizes)
    for seed in range(4):
        g = torch.Generator(device="cuda").manual_seed(seed)
        noise1 = torch.rand(n, device="cuda", generator=g)
        noise2 = torch.rand(n, device="cuda", generator=g)
        share = torch.distributions.Dirichlet(torch.ones(T)).sample().cuda() if seed else torch.tensor([1.0] + [0.0] * (T - 1)).cuda()
        same = len(set(sizes)) == 1
        mask, ids_restore, ids_keep, idx, counts, seg, slotmap, tok = K().mask_build(noise1, noise2, share, sizes, nenc, sizes[0], same)
        # reference op sequence
        want = (share * nenc).round().long()
        pre = []
        off = 0
        for t, nt in enumerate(sizes):
            order = torch.argsort(noise1[off:off + nt].unsqueeze(0), dim=1, stable=True)
            rank = torch.gather(torch.arange(nt, device="cuda").unsqueeze(0), 1, order)
            pre.append(torch.where(rank < want[t], 0, 1))
            off += nt
        flat = torch.cat(pre, dim=1)
        ids_shuffle = torch.argsort(flat + noise2.unsqueeze(0), dim=1, stable=True)
        r_restore = torch.argsort(ids_shuffle, dim=1)
        r_keep = ids_shuffle[:, :nenc]
        m = torch.ones_like(flat)
        m[:, :nenc] = 0
        m = torch.gather(m, 1, r_restore)
        assert torch.equal(mask, m[0]) and torch.equal(ids_restore, r_restore[0]) and torch.equal(ids_keep, r_keep[0])
        off, acc = 0, 0
        assert int(seg[0]) == 0
        for t, nt in enumerate(sizes):
            ix = (m[0, off:off + nt] == 0).nonzero(as_tuple=True)[0]
            assert int(counts[t]) == ix.numel()
            assert torch.equal(idx[off:off + ix.numel()].long(), ix)
            acc += ix.numel()
            assert int(seg[t + 1]) == acc
            if same:
                sm = torch.full((nt,), -1, dtype=torch.int32, device="cuda")
                sm[ix] = torch.arange(ix.numel(), dtype=torch.int32, device="cuda")
                assert torch.equal(slotmap[t], sm)
            off += nt
        # token table: the visible tokens in encoder order = the kept global ids in ascending order
        assert torch.equal(tok.long(), (m[0] == 0).nonzero(as_tuple=True)[0])
        assert acc == nenc and int(seg[T + 1]) == nenc + sizes[0]


@pytest.mark.parametrize("dtype", [f32, bf16])
@pytest.mark.parametrize("B,D", [(4, 192), (37, 768), (256, 1024)])
def test_dino_loss_fused(B, D, dtype):
    """fused forward + student gradient against the reference op sequence (criterion.py:328-335) in fp32: <= 1e-5"""
    from incomplete_multimodal_fusion_b200.multimae.criterion import dino_loss_func
    s = rnd(B, 3, D, seed=1)[:, 1].to(dtype).requires_grad_(True)      # strided rows, like the pooled return tokens
    t = rnd(B, D, seed=2).to(dtype)
    loss = dino_loss_func(s, t)
    (loss * 1.7).backward()
    s32 = s.detach().float().requires_grad_(True)
    so = F.log_softmax(F.normalize(s32, dim=1) / 0.1, dim=-1)
    to = F.softmax(F.normalize(t.float(), dim=1) / 0.04, dim=-1)
    ref = (-to * so).sum(-1).mean()
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    tol = 1e-5 if dtype == f32 else 1e-2
    assert rel(s.grad, s32.grad) < tol


def test_hard_negative_loss_matches_golden_and_oracle(golden_dir):
    """HardNegtive_loss (criterion.py:233-268) with the similarity on the tcgen05 GEMM: golden value from the reference's
    arithmetic (fp32) within the bf16 operand tolerance, gradients against the oracle"""
    import os
    import oracle
    from incomplete_multimodal_fusion_b200.multimae.criterion import HardNegtive_loss
    fx = torch.load(os.path.join(golden_dir, "losses.pt"), weights_only=False)
    g = torch.Generator().manual_seed(fx["seed"])      # same draw order as tests/golden/make_golden.py
    torch.randn(3, 2, 32, 32, generator=g); torch.randn(3, 2, 32, 32, generator=g); torch.rand(3, 16, generator=g)
    fa, fb = torch.randn(6, 48, generator=g), torch.randn(6, 48, generator=g)
    a, b = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
    loss = HardNegtive_loss()(a, b)
    assert abs(float(loss) - float(fx["hardneg"])) < 1e-2 * abs(float(fx["hardneg"]))
    loss.backward()
    a2, b2 = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
    oracle.hard_negative_loss(a2, b2).backward()
    assert rel(a.grad, a2.grad) < 3e-2 and rel(b.grad, b2.grad) < 3e-2
    from incomplete_multimodal_fusion_b200.multimae.criterion import dino_loss_func
    d = dino_loss_func(fa.cuda(), fb.cuda())
    assert abs(float(d) - float(fx["dino"])) < 1e-5 * abs(float(fx["dino"]))


@pytest.mark.parametrize("B,D,beta,tau_plus,temp", [(64, 384, 1.0, 0.1, 0.5), (100, 768, 0.5, 0.05, 0.2), (2, 40, 1.0, 0.1, 0.5)])
def test_hard_negative_loss_fused_kernel(B, D, beta, tau_plus, temp):
    """mmf_hardneg_loss (forward + both input gradients in one call) against the reference's arithmetic in fp32 autograd
    (oracle restatement of criterion.py:233-268) at pre-training batch sizes; also the reference class itself when
    baseline/_ref is available (its .cuda() call runs here), and the 'easy' estimator against autograd of its formula"""
    import oracle
    from incomplete_multimodal_fusion_b200.multimae.criterion import HardNegtive_loss
    g = torch.Generator().manual_seed(5)
    fa, fb = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g) * 0.5 + 0.3 * torch.randn(B, D, generator=torch.Generator().manual_seed(5))
    a, b = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
    loss = HardNegtive_loss(tau_plus=tau_plus, beta=beta, temperature=temp)(a, b)
    (loss * 1.7).backward()
    a2, b2 = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
    ref = oracle.hard_negative_loss(a2, b2, tau_plus=tau_plus, beta=beta, temperature=temp)
    (ref * 1.7).backward()
    # the similarity takes bf16-rounded operands (torch.mm under the reference's autocast): 1e-2 on the loss, 3e-2 on gradients
    assert abs(float(loss) - float(ref)) < 1e-2 * abs(float(ref)), (float(loss), float(ref))
    assert rel(a.grad, a2.grad) < 3e-2 and rel(b.grad, b2.grad) < 3e-2, (rel(a.grad, a2.grad), rel(b.grad, b2.grad))
    try:
        from baseline import harness as H
        refmod = H.load().criterion.HardNegtive_loss(tau_plus=tau_plus, beta=beta, temperature=temp) if H.available() else None
    except Exception:
        refmod = None
    if refmod is not None:
        a3, b3 = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
        r3 = refmod(a3, b3)
        r3.backward()
        assert abs(float(loss) - float(r3)) < 1e-2 * abs(float(r3))
        assert rel(a.grad / 1.7, a3.grad) < 3e-2
    # 'easy' estimator: Ng = sum of the negatives
    a4, b4 = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
    le = HardNegtive_loss(temperature=temp, estimator='easy')(a4, b4)
    le.backward()
    a5, b5 = fa.cuda().requires_grad_(True), fb.cuda().requires_grad_(True)
    o1, o2 = torch.nn.functional.normalize(a5, dim=1), torch.nn.functional.normalize(b5, dim=1)
    out = torch.cat([o1, o2], 0)
    neg = torch.exp(out @ out.t() / temp)
    eye = torch.eye(B, dtype=torch.bool, device="cuda")
    keep = ~torch.cat([torch.cat([eye, eye], 1), torch.cat([eye, eye], 1)], 0)
    Ng = (neg * keep).sum(-1)
    pos = torch.exp((o1 * o2).sum(-1) / temp)
    pos = torch.cat([pos, pos], 0)
    lr = (-torch.log(pos / (pos + Ng))).mean()
    lr.backward()
    assert abs(float(le) - float(lr)) < 1e-2 * abs(float(lr))
    assert rel(a4.grad, a5.grad) < 3e-2 and rel(b4.grad, b5.grad) < 3e-2


@pytest.mark.parametrize("sizes,keep", [((49, 49, 49), (10, 0, 49)), ((196, 196, 196), (196, 0, 0)), ((64, 64), (1, 63))])
def test_mask_explicit_matches_torch_ops(sizes, keep):
    """caller-provided masks (multimae.py:372-376, 378-383): stable argsort of the 0 / 1 row, ids_restore, ids_keep, the
    per-task nonzero() lists, counts, segment table, slot map and token table, all in one launch with no host sync; a wrong
    num_encoded_tokens sets the error flag and still yields tables that describe exactly nenc tokens"""
    g = torch.Generator().manual_seed(3)
    rows = []
    for nt, k in zip(sizes, keep):
        m = torch.ones(nt, dtype=torch.int64)
        m[torch.randperm(nt, generator=g)[:k]] = 0
        rows.append(m)
    given = torch.cat(rows).cuda()
    nenc = sum(keep)
    same = len(set(sizes)) == 1
    ids_restore, ids_keep, idx, counts, seg, slotmap, tok, err = K().mask_explicit(given, sizes, nenc, sizes[0], same)
    assert int(err) == 0
    shuf = torch.argsort(given.unsqueeze(0), dim=1, stable=True)
    assert torch.equal(ids_restore, torch.argsort(shuf, dim=1)[0]) and torch.equal(ids_keep, shuf[0, :nenc])
    assert torch.equal(tok.long(), (given == 0).nonzero(as_tuple=True)[0])
    off, acc = 0, 0
    for t, nt in enumerate(sizes):
        ix = (given[off:off + nt] == 0).nonzero(as_tuple=True)[0]
        assert int(counts[t]) == ix.numel() and torch.equal(idx[off:off + ix.numel()].long(), ix)
        acc += ix.numel()
        assert int(seg[t + 1]) == acc
        if same:
            sm = torch.full((nt,), -1, dtype=torch.int32, device="cuda")
            sm[ix] = torch.arange(ix.numel(), dtype=torch.int32, device="cuda")
            assert torch.equal(slotmap[t], sm)
        off += nt
    assert int(seg[len(sizes) + 1]) == nenc + sizes[0]
    for wrong in (nenc - 1, nenc + 3):
        if 0 < wrong <= sum(sizes):
            r = K().mask_explicit(given, sizes, wrong, sizes[0], same)
            assert int(r[-1]) == 1 and int(r[3].sum()) == wrong and int(r[4][len(sizes)]) == wrong


def test_im2col_tokens_matches_per_modality_gather():
    """the token-table im2col over all modalities against the per-modality visible-patch gather it replaces: each row holds
    its token's patch in its modality's column block, the one-hot modality flag, and zeros elsewhere"""
    B, P, H = 3, 8, 32
    chans = (1, 3, 1)
    g = torch.Generator().manual_seed(0)
    imgs = [torch.randn(B, c, H, H, generator=g).cuda() for c in chans]
    F_ = (H // P) ** 2
    keep = [torch.tensor(sorted(torch.randperm(F_, generator=g)[:k].tolist()), dtype=torch.int32) for k in (5, 0, 9)]
    tok = torch.cat([k + m * F_ for m, k in enumerate(keep)]).to(torch.int32).cuda()
    nenc = tok.numel()
    ks = [c * P * P for c in chans]
    col_off = [0, ks[0], ks[0] + ks[1]]
    ktot = sum(ks)
    A = torch.full((B * nenc, ktot + 8), float("nan"), dtype=bf16, device="cuda")
    K().im2col_tokens(imgs, tok, A, P, col_off, [0, F_, 2 * F_, 3 * F_], ktot)
    want = torch.zeros(B, nenc, ktot + 8, dtype=bf16, device="cuda")
    o = 0
    for m, k in enumerate(keep):
        if k.numel():
            ref = torch.empty(B * k.numel(), ks[m], dtype=bf16, device="cuda")
            K().im2col_gather(imgs[m], k.cuda(), ref, P)
            want[:, o:o + k.numel(), col_off[m]:col_off[m] + ks[m]] = ref.view(B, k.numel(), ks[m])
            want[:, o:o + k.numel(), ktot + m] = 1
        o += k.numel()
    assert torch.equal(A.view(B, nenc, -1), want)


@pytest.mark.parametrize("kind", ["mse", "l1"])
def test_masked_loss_norm_pix(kind):
    """norm_pix=True (criterion.py:90-96, 147-153): the target standardised per patch; against the reference class itself
    when baseline/_ref is present, else against the same arithmetic restated in torch"""
    from incomplete_multimodal_fusion_b200.multimae.criterion import MaskedL1Loss, MaskedMSELoss
    B, C, H, P = 3, 2, 32, 8
    g = torch.Generator().manual_seed(4)
    pred = torch.randn(B, C, H, H, generator=g).cuda().requires_grad_(True)
    tgt = (torch.randn(B, C, H, H, generator=g) * 3 + 1).cuda()
    mask = (torch.rand(B, (H // P) ** 2, generator=g) > 0.4).long().cuda()
    ours = (MaskedMSELoss if kind == "mse" else MaskedL1Loss)(patch_size=P, norm_pix=True)
    loss = ours(pred, tgt, mask=mask)
    loss.backward()
    ref_cls = None
    try:
        from baseline import harness as H_
        if H_.available():
            crit = H_.load().criterion
            ref_cls = crit.MaskedMSELoss if kind == "mse" else crit.MaskedL1Loss
    except Exception:
        ref_cls = None
    p2 = pred.detach().clone().requires_grad_(True)
    if ref_cls is not None:
        ref = ref_cls(patch_size=P, norm_pix=True)(p2, tgt, mask=mask)
    else:
        t = tgt.view(B, C, H // P, P, H // P, P)
        t = ((t - t.mean(dim=(1, 3, 5), keepdim=True)) / torch.sqrt(t.var(dim=(1, 3, 5), keepdim=True) + 1e-6)).view(B, C, H, H)
        e = (p2 - t) ** 2 if kind == "mse" else (p2 - t).abs()
        m = mask.view(B, H // P, H // P).repeat_interleave(P, 1).repeat_interleave(P, 2).float()
        ref = ((e.mean(1) * m).flatten(1).sum(1) / m.flatten(1).sum(1)).nanmean()
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel(pred.grad, p2.grad) < 1e-5
