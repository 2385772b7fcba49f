"""Kernel-level parity (-m gpu): every libc2d entry point, called through the C ABI (ctypes), against a plain
PyTorch fp32 evaluation of the same op (tests/torch_ops.py) on identical inputs.
Tolerances: fp32 kernels 2e-5 rel-L2 (summation order only); bf16 kernels 1e-2 (inputs are rounded to bf16
first, so the only differences are fp32-accumulate order and the final bf16 rounding)."""
import numpy as np
import pytest
import torch

import torch_ops as T
from clap2diffusion_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
F32, BF16 = torch.float32, torch.bfloat16


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, dtype=F32, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(hash((shape, seed)) % (2 ** 31))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def tol(dtype):
    return 2e-5 if dtype == F32 else 1e-2


# ------------------------------------------------------------------ linear
LIN_CASES = [
    # M, N, K, impl, dtype
    (300, 200, 96, ops.IMPL_SIMT, F32), (5, 24576, 256, ops.IMPL_SIMT, F32), (77, 3, 6, ops.IMPL_SIMT, F32),
    (130, 70, 36, ops.IMPL_SIMT, F32), (256, 320, 320, ops.IMPL_SIMT, BF16),
    (8192, 320, 320, ops.IMPL_TCGEN05, BF16), (2048, 640, 2560, ops.IMPL_TCGEN05, BF16),
    (512, 1280, 1280, ops.IMPL_TCGEN05, BF16), (128, 1280, 5120, ops.IMPL_TCGEN05, BF16),
    (1000, 200, 72, ops.IMPL_TCGEN05, BF16), (154, 1536, 768, ops.IMPL_TCGEN05, BF16),
    (4096, 4096, 512, ops.IMPL_TCGEN05, BF16), (64, 8, 64, ops.IMPL_TCGEN05, BF16),
]


@pytest.mark.parametrize("M,N,K,impl,dtype", LIN_CASES)
def test_linear(M, N, K, impl, dtype):
    x, w = rnd(M, K, dtype=dtype), rnd(N, K, dtype=dtype, scale=K ** -0.5, seed=1)
    b, r = rnd(N, seed=2), rnd(M, N, dtype=dtype, seed=3)
    for act in (ops.ACT_NONE, ops.ACT_GELU, ops.ACT_SILU):
        y = ops.linear(x, w, b, act=act, residual=r, impl=impl)
        assert rel(y, T.linear(x, w, b, act=act, residual=r)) < tol(dtype), (act,)
    y = ops.linear(x, w, impl=impl)
    assert rel(y, T.linear(x, w)) < tol(dtype)


@pytest.mark.parametrize("impl,dtype", [(ops.IMPL_SIMT, F32), (ops.IMPL_TCGEN05, BF16)])
def test_linear_strided_and_rowvec(impl, dtype):
    B, N_, C = 2, 256, 320
    qkv = rnd(B, N_, 3 * C, dtype=dtype)
    w = rnd(C, C, dtype=dtype, scale=C ** -0.5, seed=1)
    x = qkv[..., C:2 * C]                       # strided rows (ldx = 3C)
    y = ops.linear(x, w, impl=impl)
    assert rel(y, T.linear(x, w)) < tol(dtype)
    out = torch.zeros(B, N_, 2 * C, device=DEV, dtype=dtype)
    ops.linear(x, w, out=out[..., C:], impl=impl)                                   # strided output (ldy = 2C)
    assert rel(out[..., C:], T.linear(x, w)) < tol(dtype) and float(out[..., :C].abs().max()) == 0.0
    rv = rnd(B, C, seed=5)
    y = ops.linear(x, w, rowvec=rv, rows_per_vec=N_, impl=impl)
    assert rel(y, T.linear(x, w, rowvec=rv, rows_per_vec=N_)) < tol(dtype)


def test_geglu_fused_vs_unfused():
    M, C = 1024, 320
    x = rnd(M, C, dtype=BF16)
    w, b = rnd(8 * C, C, scale=C ** -0.5, seed=1), rnd(8 * C, seed=2, scale=0.1)
    wp, bp = ops.pack_geglu(w, b, BF16)
    wp_ref, bp_ref = T.pack_geglu(w, b, BF16)
    assert torch.equal(wp, wp_ref) and torch.equal(bp, bp_ref)
    y = ops.geglu_linear(x, wp, bp)
    ref = T.geglu(T.linear(x.float(), w.to(BF16).float(), b))
    assert rel(y, ref) < 1e-2
    y32 = ops.geglu(ops.linear(x.float(), w, b, impl=ops.IMPL_SIMT))
    assert rel(y32, T.geglu(T.linear(x.float(), w, b))) < 2e-5


# ------------------------------------------------------------------ conv3x3
CONV_SIMT = [(2, 16, 16, 4, 320, 1, False), (1, 16, 16, 64, 96, 2, False), (2, 8, 8, 128, 64, 1, True),
             (1, 9, 7, 24, 40, 1, False), (1, 9, 7, 24, 40, 2, False)]


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,up", CONV_SIMT)
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_conv3x3_simt(B, H, W, Cin, Cout, stride, up, dtype):
    x = rnd(B, H, W, Cin, dtype=dtype)
    w = ops.pack_conv3x3(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1), dtype)
    b = rnd(Cout, seed=2)
    y = ops.conv3x3(x, w, b, stride=stride, upsample=up, impl=ops.IMPL_SIMT)
    ref = T.conv3x3(x, w, b, stride=stride, upsample=up)
    assert y.shape == ref.shape
    assert rel(y, ref) < tol(dtype)


CONV_TC = [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280), (2, 8, 8, 1280, 1280),
           (3, 8, 8, 2560, 1280), (1, 32, 32, 960, 640), (1, 64, 64, 320, 4), (1, 128, 128, 128, 128),
           (1, 256, 256, 64, 64), (5, 4, 4, 64, 32), (1, 16, 16, 72, 48)]


@pytest.mark.parametrize("B,H,W,Cin,Cout", CONV_TC)
def test_conv3x3_tcgen05(B, H, W, Cin, Cout):
    x = rnd(B, H, W, Cin, dtype=BF16)
    w = ops.pack_conv3x3(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1), BF16)
    b, rv, r = rnd(Cout, seed=2), rnd(B, Cout, seed=3), rnd(B, H, W, Cout, dtype=BF16, seed=4)
    y = ops.conv3x3(x, w, b, rowvec=rv, residual=r, impl=ops.IMPL_TCGEN05)
    assert rel(y, T.conv3x3(x, w, b, rowvec=rv, residual=r)) < 1e-2
    y2 = ops.conv3x3(x, w, b, impl=ops.IMPL_TCGEN05)
    ys = ops.conv3x3(x, w, b, impl=ops.IMPL_SIMT)
    assert rel(y2, T.conv3x3(x, w, b)) < 1e-2
    assert rel(y2, ys) < 1e-2            # tcgen05 vs FFMA on identical bf16 inputs


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (3, 16, 16, 1280, 1280), (1, 8, 8, 64, 32)])
def test_conv3x3_stride2_tcgen05(B, H, W, Cin, Cout):
    """Downsample2D conv: stride-2 gather through TMA element strides."""
    x = rnd(B, H, W, Cin, dtype=BF16)
    w = ops.pack_conv3x3(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1), BF16)
    b = rnd(Cout, seed=2)
    y = ops.conv3x3(x, w, b, stride=2, impl=ops.IMPL_TCGEN05)
    ref = T.conv3x3(x, w, b, stride=2)
    assert y.shape == ref.shape and rel(y, ref) < 1e-2


def test_pack_conv():
    w = rnd(24, 16, 3, 3)
    assert torch.equal(ops.pack_conv3x3(w, F32), T.pack_conv3x3(w, F32))


# ------------------------------------------------------------------ norms
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("B,HW,C", [(2, 4096, 320), (2, 64, 1280), (3, 256, 2560), (1, 1024, 1920), (1, 16384, 128)])
def test_group_norm(B, HW, C, dtype):
    x = rnd(B, HW, C, dtype=dtype) * 2 + 0.5
    g, b = 1 + 0.1 * rnd(C, seed=1), 0.1 * rnd(C, seed=2)
    for silu in (False, True):
        y = ops.group_norm(x, g, b, 32, 1e-5, silu)
        assert rel(y, T.group_norm(x, g, b, 32, 1e-5, silu)) < (5e-5 if dtype == F32 else 1e-2)


def _ref_stats(y, B):
    """fp64 per-channel (sum, sumsq) of y [B, ..., C] in the fixed-point units of the epilogue."""
    C = y.shape[-1]
    yy = y.double().reshape(B, -1, C)
    return torch.stack([yy.sum(1), (yy * yy).sum(1)], -1) * T.STATS_SCALE


def _stats_close(stats, y, B, C):
    got, ref = stats.view(B, C, 2).double(), _ref_stats(y, B)
    # sums: absolute error relative to the sum of |y|; sums of squares: relative
    assert float((got[..., 1] - ref[..., 1]).abs().max() / ref[..., 1].abs().max()) < 2e-3
    scale = y.double().abs().reshape(B, -1, C).sum(1) * T.STATS_SCALE
    assert float(((got[..., 0] - ref[..., 0]).abs() / scale).max()) < 2e-3


@pytest.mark.parametrize("B,HW,N,K,K1,res", [(2, 4096, 320, 320, 0, True), (2, 1024, 640, 1920, 1280, False),
                                             (4, 64, 1280, 2560, 1280, True), (3, 256, 200, 72, 0, False),
                                             (2, 1024, 640, 960, 640, True), (2, 16, 1280, 1280, 0, True),
                                             (2, 4, 1280, 2560, 1280, False)])
def test_linear_ex_stats_and_kconcat(B, HW, N, K, K1, res):
    """c2d_linear_ex: A = [x | x2] along K, epilogue channel statistics == statistics of the stored output."""
    M = B * HW
    x, w = rnd(M, K, dtype=BF16), rnd(N, K, dtype=BF16, scale=K ** -0.5, seed=1)
    b = rnd(N, seed=2)
    r = rnd(M, N, dtype=BF16, seed=3) if res else None
    stats = torch.zeros(B * N * 2, device=DEV, dtype=torch.int64)
    if K1:
        xa, xb = x[:, :K1].contiguous(), x[:, K1:].contiguous()
        y = ops.linear(xa, w, b, residual=r, x2=xb, stats=stats, stats_rows=HW)
    else:
        y = ops.linear(x, w, b, residual=r, stats=stats, stats_rows=HW)
    ref = T.linear(x, w, b, residual=r)
    assert rel(y, ref) < 1e-2
    assert torch.equal(y, ops.linear(x, w, b, residual=r, impl=ops.IMPL_TCGEN05))   # same arithmetic as the plain entry point
    _stats_close(stats, y, B, N)
    # bit-reproducible: integer accumulation does not depend on CTA order
    stats2 = torch.zeros_like(stats)
    if K1:
        ops.linear(xa, w, b, residual=r, x2=xb, stats=stats2, stats_rows=HW)
    else:
        ops.linear(x, w, b, residual=r, stats=stats2, stats_rows=HW)
    assert torch.equal(stats, stats2)


@pytest.mark.parametrize("B,H,Cin,Cout,stride", [(2, 32, 320, 640, 1), (2, 16, 640, 320, 1), (4, 8, 1280, 1280, 1),
                                                 (2, 32, 320, 320, 2), (1, 64, 128, 96, 1), (2, 4, 1280, 1280, 1), (2, 2, 1280, 1280, 1), (2, 4, 640, 640, 2),
                                                 (2, 64, 8, 320, 1), (1, 16, 24, 64, 1),
                                                 # benchmark-batch 8x8 planes: the split-K path of the CTA-pair kernel
                                                 (16, 8, 1280, 1280, 1), (16, 8, 2560, 1280, 1), (16, 16, 1280, 1280, 2)])
def test_conv3x3_ex_stats(B, H, Cin, Cout, stride):
    x = rnd(B, H, H, Cin, dtype=BF16)
    w = ops.pack_conv3x3(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1), BF16)
    b, rv = rnd(Cout, seed=2), rnd(B, Cout, seed=4)
    Ho = H // stride
    r = rnd(B, Ho, Ho, Cout, dtype=BF16, seed=3)
    stats = torch.zeros(B * Cout * 2, device=DEV, dtype=torch.int64)
    y = ops.conv3x3(x, w, b, rowvec=rv, residual=r, stride=stride, stats=stats)
    assert torch.equal(y, ops.conv3x3(x, w, b, rowvec=rv, residual=r, stride=stride, impl=ops.IMPL_TCGEN05))
    assert rel(y, T.conv3x3(x, w, b, rowvec=rv, residual=r, stride=stride)) < 1e-2
    _stats_close(stats, y, B, Cout)


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride", [(2, 96, 64, 320, 320, 1), (1, 48, 32, 640, 640, 1), (2, 24, 16, 1280, 1280, 1),
                                                   (2, 12, 8, 1280, 1280, 1), (2, 96, 64, 320, 320, 2), (1, 6, 4, 640, 640, 1),
                                                   # output widths the tcgen05 tiler does not take -> FFMA kernel + stand-alone statistics
                                                   (2, 64, 96, 320, 320, 1), (1, 12, 24, 640, 320, 1), (2, 64, 96, 320, 320, 2), (1, 16, 16, 20, 32, 1)])
def test_conv3x3_non_square_planes(B, H, W, Cin, Cout, stride):
    """Latents of non-square images (e.g. 768 x 512 -> 96 x 64): heights that are not powers of two run on the tcgen05
    tiler, widths it cannot box fall back to the FFMA kernel inside c2d_conv3x3_ex -- both with channel statistics."""
    x = rnd(B, H, W, Cin, dtype=BF16)
    w = ops.pack_conv3x3(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1), BF16)
    b, rv = rnd(Cout, seed=2), rnd(B, Cout, seed=4)
    Ho, Wo = H // stride, W // stride
    r = rnd(B, Ho, Wo, Cout, dtype=BF16, seed=3)
    stats = torch.zeros(B * Cout * 2, device=DEV, dtype=torch.int64)
    y = ops.conv3x3(x, w, b, rowvec=rv, residual=r, stride=stride, stats=stats)
    assert tuple(y.shape) == (B, Ho, Wo, Cout)
    assert rel(y, T.conv3x3(x, w, b, rowvec=rv, residual=r, stride=stride)) < 1e-2
    assert rel(y, ops.conv3x3(x, w, b, rowvec=rv, residual=r, stride=stride, impl=ops.IMPL_SIMT)) < 1e-2
    if Cout % 8 == 0:
        _stats_close(stats, y, B, Cout)


@pytest.mark.parametrize("B,HW,C1,C2", [(2, 4096, 320, 0), (2, 64, 1280, 1280), (3, 256, 1280, 640), (1, 1024, 640, 320),
                                        (16, 1024, 640, 0), (2, 4096, 640, 320), (1, 16384, 128, 0)])
def test_group_norm_apply_from_channel_stats(B, HW, C1, C2):
    x1 = rnd(B, HW, C1, dtype=BF16) * 2 + 0.5
    x2 = rnd(B, HW, C2, dtype=BF16, seed=7) * 3 - 1 if C2 else None
    C = C1 + C2
    g, b = 1 + 0.1 * rnd(C, seed=1), 0.1 * rnd(C, seed=2)
    s1 = ops.channel_stats(x1, torch.zeros(B * C1 * 2, device=DEV, dtype=torch.int64))
    _stats_close(s1, x1, B, C1)
    s2 = ops.channel_stats(x2, torch.zeros(B * C2 * 2, device=DEV, dtype=torch.int64)) if C2 else None
    for silu in (False, True):
        y = ops.group_norm_apply(x1, s1, g, b, 32, 1e-5, silu, x2=x2, stats2=s2)
        assert rel(y, T.group_norm(x1, g, b, 32, 1e-5, silu, x2=x2)) < 1e-2
        assert rel(y, ops.group_norm(x1, g, b, 32, 1e-5, silu, x2=x2)) < 4e-3      # vs the two-pass kernel


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_group_norm_two_source(dtype):
    B, HW, C1, C2 = 2, 1024, 640, 320
    x1, x2 = rnd(B, HW, C1, dtype=dtype), rnd(B, HW, C2, dtype=dtype, seed=7) * 3
    g, b = 1 + 0.1 * rnd(C1 + C2, seed=1), 0.1 * rnd(C1 + C2, seed=2)
    raw = torch.empty(B, HW, C1 + C2, device=DEV, dtype=dtype)
    y = ops.group_norm(x1, g, b, 32, 1e-5, True, x2=x2, raw_cat=raw)
    assert torch.equal(raw, torch.cat([x1, x2], -1))
    assert rel(y, T.group_norm(x1, g, b, 32, 1e-5, True, x2=x2)) < (5e-5 if dtype == F32 else 1e-2)
    assert torch.equal(ops.concat(x1, x2), torch.cat([x1, x2], -1))


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("M,C", [(8192, 320), (300, 1280), (77, 768), (5, 6), (33, 192), (4099, 96), (1000, 64), (37, 24), (513, 128)])
def test_layer_norm(M, C, dtype):
    x = rnd(M, C, dtype=dtype) * 3 + 1
    g, b = 1 + 0.1 * rnd(C, seed=1), 0.1 * rnd(C, seed=2)
    assert rel(ops.layer_norm(x, g, b, 1e-5), T.layer_norm(x, g, b, 1e-5)) < (2e-5 if dtype == F32 else 1e-2)


# ------------------------------------------------------------------ attention
ATT = [(2, 8, 1024, 1024, 40), (1, 8, 256, 256, 160), (2, 8, 64, 64, 160), (2, 8, 1024, 77, 80), (2, 8, 300, 81, 40),
       (3, 4, 10, 10, 48), (2, 8, 77, 10, 32), (2, 8, 16, 16, 96), (2, 8, 256, 16, 64)]


@pytest.mark.parametrize("B,h,Nq,Nkv,d", ATT)
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_attention_simt(B, h, Nq, Nkv, d, dtype):
    q, k, v = rnd(B, Nq, h * d, dtype=dtype), rnd(B, Nkv, h * d, dtype=dtype, seed=1), rnd(B, Nkv, h * d, dtype=dtype, seed=2)
    o = ops.attention(q, k, v, h, impl=ops.IMPL_SIMT)
    assert rel(o, T.attention(q, k, v, h)) < (2e-5 if dtype == F32 else 1e-2)


ATT_TC = [(2, 8, 4096, 4096, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160), (3, 8, 64, 64, 160),
          (2, 8, 4096, 77, 40), (2, 8, 1024, 81, 80), (1, 8, 256, 77, 160), (2, 4, 200, 333, 64), (1, 2, 130, 129, 16),
          (1, 3, 700, 1000, 48), (2, 2, 256, 256, 24), (1, 1, 1, 257, 64)]


@pytest.mark.parametrize("B,h,Nq,Nkv,d", ATT_TC)
def test_attention_tcgen05(B, h, Nq, Nkv, d):
    C = h * d
    if Nq == Nkv:        # self-attention: views of one packed QKV buffer, as the UNet uses it
        qkv = rnd(B, Nq, 3 * C, dtype=BF16)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:                # cross-attention: q dense, K/V views of the cached [B, T, 2C] buffer
        q = rnd(B, Nq, C, dtype=BF16)
        kv = rnd(B, Nkv, 2 * C, dtype=BF16, seed=1)
        k, v = kv[..., :C], kv[..., C:]
    o = ops.attention(q, k, v, h, impl=ops.IMPL_TCGEN05)
    ref = T.attention(q, k, v, h)
    assert rel(o, ref) < 1e-2
    assert rel(o, ops.attention(q, k, v, h, impl=ops.IMPL_SIMT)) < 1e-2
    # large-magnitude logits exercise the lazy-rescale path
    o2 = ops.attention(q * 6, k, v, h, impl=ops.IMPL_TCGEN05)
    assert rel(o2, T.attention(q * 6, k, v, h)) < 2e-2


def test_attention_packed_views_mask_and_small():
    B, N_, C, h = 2, 256, 320, 8
    qkv = rnd(B, N_, 3 * C)
    o = ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], h, impl=ops.IMPL_SIMT)
    assert rel(o, T.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], h)) < 2e-5
    # key-padding mask (AudioCrossAttention semantics)
    q, k, v = rnd(B, 64, 512), rnd(B, 16, 512, seed=1), rnd(B, 16, 512, seed=2)
    mask = torch.ones(B, 16, dtype=torch.bool, device=DEV)
    mask[:, 12:] = False
    assert rel(ops.attention(q, k, v, 8, mask=mask), T.attention(q, k, v, 8, mask=mask)) < 2e-5
    # single head d = 768 with batch-broadcast queries (AudioTokenGenerator)
    q0 = rnd(1, 16, 768)
    kv = rnd(B, 16, 2, 768, seed=3)
    o = ops.attention(q0.expand(B, 16, 768), kv[:, :, 0], kv[:, :, 1], 1)
    assert rel(o, T.attention(q0.expand(B, 16, 768), kv[:, :, 0], kv[:, :, 1], 1)) < 2e-5


# ------------------------------------------------------------------ audio context + small kernels
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("mode", [ops.AUDIO_ADD, ops.AUDIO_CONCAT])
@pytest.mark.parametrize("K", [10, 16, 3])
def test_audio_context(dtype, mode, K):
    B = 2
    ehs, audio = rnd(B, 77, 768, dtype=dtype), rnd(B, K, 768, dtype=dtype, seed=1) * 0.3
    w1, b1 = rnd(64, 768, dtype=dtype, scale=768 ** -0.5, seed=2), rnd(64, seed=3, scale=0.1)
    w2, b2 = rnd(768, 64, dtype=dtype, scale=0.125, seed=4), rnd(768, seed=5, scale=0.1)
    alpha = torch.tensor([0.3], device=DEV)
    y = ops.audio_context(ehs, audio, w1, b1, w2, b2, alpha, mode)
    ref = T.audio_context(ehs, audio, w1, b1, w2, b2, alpha, mode)
    assert y.shape == ref.shape and rel(y, ref) < (1e-5 if dtype == F32 else 1e-2)


def test_elementwise_and_layout():
    t = torch.tensor([981.0, 1.0, 500.0], device=DEV)
    # angles reach ~1e3 rad: a 1-ulp difference in the fp32 frequency moves cos/sin by ~6e-5
    assert rel(ops.timestep_embedding(t, 320), T.timestep_embedding(t, 320)) < 2e-5
    x = rnd(3, 1000, 640)
    assert rel(ops.geglu(x), T.geglu(x)) < 1e-6
    assert rel(ops.unary(x, ops.ACT_SILU), T.unary(x, 2)) < 1e-6
    assert torch.equal(ops.cast(x, BF16), x.to(BF16))
    assert rel(ops.add(x, x * 2), x * 3) < 1e-7
    y = rnd(2, 8, 16, 64)
    assert torch.equal(ops.upsample2x(y), T.upsample2x(y))
    z = rnd(2, 4, 16, 24)
    nh = ops.nchw_to_nhwc(z, F32)
    assert torch.equal(nh, z.permute(0, 2, 3, 1).contiguous()) and torch.equal(ops.nhwc_to_nchw(nh), z)
    m = rnd(5, 37, 91)
    assert torch.equal(ops.transpose(m), m.transpose(1, 2).contiguous())
    s = rnd(300, 4096)
    assert rel(ops.softmax_rows(s, 0.044), T.softmax_rows(s, 0.044)) < 1e-5


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_cfg_sched_step(dtype):
    B, H, W = 3, 16, 16
    eps2 = rnd(2 * B, H, W, 4, dtype=dtype)
    x = rnd(B, 4, H, W, seed=1)
    coef = torch.tensor([0.97, -0.12, 0.8], device=DEV)
    x_ref, xin_ref = x.clone(), torch.empty(2 * B, H, W, 4, device=DEV, dtype=dtype)
    T.cfg_sched_step(eps2, x_ref, xin_ref, 7.5, coef)
    xin, tr = torch.empty_like(xin_ref), torch.empty_like(x)
    ops.cfg_sched_step(eps2, x, xin, 7.5, coef, trace=tr)
    assert rel(x, x_ref) < 1e-6 and torch.equal(tr, x) and rel(xin, xin_ref) < (1e-6 if dtype == F32 else 5e-3)


def test_audio_small_kernels():
    B, K, D = 3, 10, 768
    a, b = rnd(B, D), rnd(K, D, seed=1)
    assert rel(ops.bcast_add(a, b, B, K, D, 1, 2), T.bcast_add(a, b, B, K, D, 1, 2)) < 1e-7
    tok = rnd(B, K, D)
    anchors, w1, b1, w2, b2 = rnd(3, D, scale=0.02, seed=2), rnd(10, D, scale=0.03, seed=3), rnd(10, seed=4), rnd(3, 10, seed=5), rnd(3, seed=6)
    temp = torch.tensor(2.0, device=DEV)
    asg = ops.hier_assign(tok, anchors, w1, b1, w2, b2, temp.reshape(1))
    assert rel(asg, T.hier_assign(tok, anchors, w1, b1, w2, b2, temp)) < 1e-5
    hw = torch.softmax(rnd(B, 3, seed=7), -1)
    routing, gates = rnd(3, 3, seed=8), rnd(3, seed=9)
    for got, want in zip(ops.hier_route(tok, asg, hw, routing, gates), T.hier_route(tok, asg, hw, routing, gates)):
        assert rel(got, want) < 1e-5
    for per_sample in (False, True):
        assert rel(ops.norm_scale(tok, 60.0, per_sample), T.norm_scale(tok, 60.0, per_sample)) < 1e-5
    fg, bg, am = rnd(B, 5 * D), rnd(B, 3 * D, seed=1), rnd(B, 2 * D, seed=2)
    hwts = torch.tensor([0.5, 0.3, 0.2], device=DEV)
    assert rel(ops.legacy_combine(fg, bg, am, hwts, D), T.legacy_combine(fg, bg, am, hwts, D)) < 1e-6


def test_errors_are_loud():
    from clap2diffusion_b200._lib import C2DError
    x = rnd(4, 8)
    with pytest.raises(C2DError):
        ops.linear(x.cpu(), x.cpu())
    with pytest.raises(C2DError):   # tcgen05 path cannot take fp32
        ops.linear(x, rnd(8, 8), impl=ops.IMPL_TCGEN05)
    with pytest.raises(C2DError):
        ops.conv3x3(rnd(1, 7, 7, 8, dtype=BF16), rnd(8, 3, 3, 8, dtype=BF16), impl=ops.IMPL_TCGEN05)


# ------------------------------------------------------------------ folded LayerNorm
@pytest.mark.parametrize("M,C,N", [(4096, 320, 960), (1024, 640, 640), (256, 1280, 3840), (130, 320, 320)])
def test_layernorm_folded_into_linear(M, C, N):
    """row_stats from a producer GEMM + pack_lnfold weights: linear(ln=...) == linear(layer_norm(x))."""
    x0, w0 = rnd(M, C, dtype=BF16), rnd(C, C, dtype=BF16, scale=C ** -0.5, seed=1)
    r = rnd(M, C, dtype=BF16, seed=2) * 2 + 0.3
    rs = torch.zeros(M * 2, device=DEV, dtype=torch.int64)
    h = ops.linear(x0, w0, rnd(C, seed=3), residual=r, row_stats=rs)            # producer: stream + its row statistics
    hf = h.double()
    ref_st = torch.stack([hf.sum(1), (hf * hf).sum(1)], -1) * T.STATS_SCALE
    assert float((rs.view(M, 2).double() - ref_st).abs().max() / ref_st.abs().max()) < 2e-3
    gamma, beta = 1 + 0.2 * rnd(C, seed=4), 0.2 * rnd(C, seed=5)
    w, b = rnd(N, C, scale=C ** -0.5, seed=6), rnd(N, seed=7)
    ln = ops.pack_lnfold(w, gamma, beta, b, BF16)
    y = ops.linear(h, None, ln=ln, ln_stats=rs)
    ref = T.linear(T.layer_norm(h, gamma, beta), w.to(BF16), b)
    assert rel(y, ref) < 1e-2
    assert rel(y, ops.linear(ops.layer_norm(h, gamma, beta), w.to(BF16), b)) < 1e-2


def test_layernorm_folded_into_geglu():
    M, C, Fh = 2048, 320, 1280
    h = rnd(M, C, dtype=BF16) * 1.5 + 0.2
    rs = torch.zeros(M * 2, device=DEV, dtype=torch.int64)
    ident = torch.eye(C, device=DEV, dtype=BF16)
    h2 = ops.linear(h, ident, row_stats=rs)                                       # statistics of h itself
    assert torch.equal(h2, h)
    gamma, beta = 1 + 0.2 * rnd(C, seed=4), 0.2 * rnd(C, seed=5)
    w, b = rnd(2 * Fh, C, scale=C ** -0.5, seed=6), rnd(2 * Fh, seed=7)
    f = ops.pack_lnfold(w, gamma, beta, b, BF16, out_dtype=F32)
    wp, bp = ops.pack_geglu(f.w, f.bias, BF16)
    _, csp = ops.pack_geglu(f.w, f.colsum, BF16)
    y = ops.geglu_linear(h, None, None, ln=ops.LNFold(wp, csp, bp, 1e-5), ln_stats=rs)
    wq, bq = ops.pack_geglu(w, b, BF16)
    ref = ops.geglu_linear(ops.layer_norm(h, gamma, beta), wq, bq)
    assert rel(y, ref) < 1e-2
    ref32 = T.geglu(T.linear(T.layer_norm(h, gamma, beta).float(), w, b))
    assert rel(y, ref32) < 1e-2


# ------------------------------------------------------------------ fused cross-attention site
XATTN = [
    # B, Nq, C, T, T2, fold
    (2, 4096, 320, 77, 0, True), (2, 1024, 640, 77, 0, True), (2, 256, 1280, 77, 0, True), (3, 64, 1280, 77, 0, True),
    (2, 1024, 320, 81, 0, True), (2, 256, 640, 81, 0, False), (1, 128, 320, 77, 0, False),
    (2, 512, 320, 77, 10, True), (2, 256, 640, 77, 16, True), (1, 256, 1280, 81, 10, True), (2, 128, 320, 20, 4, False),
    # edges: partial row tiles (32 / 96 tokens per sample), the 7-chunk instances (96 keys + 16 audio keys) at every
    # head-dim class, very short key sets (TMEM-store masking path), single keys
    (3, 96, 320, 77, 0, True), (1, 32, 640, 96, 16, True), (2, 128, 1280, 96, 16, False), (2, 128, 320, 96, 16, True),
    (2, 256, 640, 5, 0, False), (1, 128, 320, 1, 1, False), (5, 64, 320, 81, 3, True),
]


@pytest.mark.parametrize("B,Nq,C,Tk,T2,fold", XATTN)
def test_xattn_fused(B, Nq, C, Tk, T2, fold):
    """c2d_xattn_fwd == to_q GEMM (folded LayerNorm) -> attention core, both as separate libc2d kernels and as torch fp32."""
    heads, M = 8, B * Nq
    assert ops.xattn_supported(torch.empty(B, Nq, C, device=DEV, dtype=BF16), heads, Tk, T2)
    h = (rnd(M, C, dtype=BF16) * 1.5 + 0.2).view(B, Nq, C)
    kv = rnd(B, Tk, 2 * C, dtype=BF16, seed=8)
    kv2 = rnd(B, T2, 2 * C, dtype=BF16, seed=9) if T2 else None
    w = rnd(C, C, scale=C ** -0.5, seed=6) * 2.0
    lam = 0.37
    kw = {}
    if fold:
        rs = torch.zeros(M * 2, device=DEV, dtype=torch.int64)
        ident = torch.eye(C, device=DEV, dtype=BF16)
        assert torch.equal(ops.linear(h, ident, row_stats=rs), h)              # statistics of h itself
        gamma, beta = 1 + 0.2 * rnd(C, seed=4), 0.2 * rnd(C, seed=5)
        ln = ops.pack_lnfold(w, gamma, beta, None, BF16)
        kw = dict(ln=ln, ln_stats=rs)
        q = ops.linear(h, None, ln=ln, ln_stats=rs)
        q32 = T.linear(T.layer_norm(h, gamma, beta), w.to(BF16))
    else:
        qb = rnd(C, seed=7)
        kw = dict(wq=w.to(BF16), q_bias=qb)
        q = ops.linear(h, w.to(BF16), qb)
        q32 = T.linear(h, w.to(BF16), qb)
    o = ops.xattn(h, ops.xattn_pack_kv(kv, heads, kv2), lambda2=lam, **kw)
    ref = ops.attention(q, kv[..., :C], kv[..., C:], heads).float()
    ref32 = T.attention(q32, kv[..., :C], kv[..., C:], heads).float()
    if T2:
        ref = ref + lam * ops.attention(q, kv2[..., :C], kv2[..., C:], heads).float()
        ref32 = ref32 + lam * T.attention(q32, kv2[..., :C], kv2[..., C:], heads).float()
    assert torch.isfinite(o.float()).all()
    assert rel(o, ref) < 1e-2, "vs the three-kernel libc2d composition"
    assert rel(o, ref32) < 1e-2, "vs torch fp32"


def test_xattn_strided_views_and_unsupported():
    B, Nq, C, heads = 2, 256, 320, 8
    big = rnd(B * Nq, 2 * C, dtype=BF16)
    x = big[:, :C].view(B, Nq, C) if False else big.view(B, Nq, 2 * C)[..., :C]       # row stride 2C
    kv = rnd(B, 77, 2 * C, dtype=BF16, seed=8)
    w = rnd(C, C, dtype=BF16, scale=C ** -0.5, seed=6)
    out = torch.zeros(B, Nq, 2 * C, device=DEV, dtype=BF16)
    kvp = ops.xattn_pack_kv(torch.cat([kv, kv], -1)[..., 2 * C:], heads)                 # strided [K | V] view (row stride 4C)
    ops.xattn(x, kvp, wq=w, out=out[..., C:])
    ref = T.attention(T.linear(x, w), kv[..., :C], kv[..., C:], heads)
    assert rel(out[..., C:], ref) < 1e-2 and float(out[..., :C].abs().max()) == 0.0
    # shapes outside the kernel are refused loudly, never silently routed elsewhere
    assert not ops.xattn_supported(torch.empty(1, 100, 320, device=DEV, dtype=BF16), 8, 77)
    kv1 = ops.xattn_pack_kv(kv[:1], heads)
    with pytest.raises(Exception):
        ops.xattn(torch.zeros(1, 100, 320, device=DEV, dtype=BF16), kv1, wq=w)
    with pytest.raises(Exception):
        ops.xattn(torch.zeros(1, 128, 320, device=DEV, dtype=F32), kv1, wq=w.float())
    with pytest.raises(Exception):
        ops.xattn_pack_kv(torch.zeros(1, 200, 640, device=DEV, dtype=BF16), heads)


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 64, 128, 128), (2, 32, 256, 256), (1, 16, 512, 512), (3, 8, 24, 40), (16, 16, 512, 512)])
def test_conv3x3_down_asymmetric_padding(B, H, Cin, Cout, dtype):
    """diffusers Downsample2D(padding=0): zeros on the right / bottom only (c2d_conv3x3_down), incl. the split-K path."""
    x = rnd(B, H, H, Cin, dtype=dtype)
    w = ops.pack_conv3x3(rnd(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=1), dtype)
    b = rnd(Cout, seed=2)
    y = ops.conv3x3_down(x, w, b)
    assert tuple(y.shape) == (B, H // 2, H // 2, Cout)
    assert rel(y, T.conv3x3_down(x, w, b)) < tol(dtype)
    assert rel(y, T.conv3x3(x, w, b, stride=2)) > 1e-2          # NOT the symmetric pad-1 convolution
    if dtype == BF16 and Cin % 8 == 0:
        stats = torch.zeros(B * Cout * 2, device=DEV, dtype=torch.int64)
        y2 = ops.conv3x3_down(x, w, b, stats=stats)
        assert torch.equal(y2, y)
        _stats_close(stats, y, B, Cout)


def test_xattn_full_size_properties():
    """BASELINE config-3 size (UNet batch 16, 4096 tokens, C = 320): size-independent properties of the fused kernel --
    a softmax-weighted mean of constant values is that constant, permuting the keys changes nothing, a zero-weight
    decoupled branch is the plain kernel, and the output is linear in V."""
    B, Nq, C, Tk, heads = 16, 4096, 320, 77, 8
    h = rnd(B * Nq, C, dtype=BF16).view(B, Nq, C)
    w = rnd(C, C, dtype=BF16, scale=C ** -0.5, seed=1)
    kv = rnd(B, Tk, 2 * C, dtype=BF16, seed=2)
    # (a) V = per-column constants -> O = the same constants (probabilities sum to one)
    cst = rnd(C, dtype=BF16, seed=3)
    kvc = kv.clone()
    kvc[..., C:] = cst
    o = ops.xattn(h, ops.xattn_pack_kv(kvc, heads), wq=w)
    assert float((o.float() - cst.float()).abs().max()) <= 2e-2 * float(cst.float().abs().max())
    # (b) key permutation invariance
    perm = torch.randperm(Tk, generator=torch.Generator().manual_seed(0)).to(DEV)
    o1 = ops.xattn(h, ops.xattn_pack_kv(kv, heads), wq=w)
    o2 = ops.xattn(h, ops.xattn_pack_kv(kv[:, perm].contiguous(), heads), wq=w)
    assert rel(o2, o1) < 5e-3
    # (c) a decoupled branch with weight 0 changes nothing (the 6-chunk instance sums the row in a different order: rounding only)
    kv2 = rnd(B, 10, 2 * C, dtype=BF16, seed=4)
    o3 = ops.xattn(h, ops.xattn_pack_kv(kv, heads, kv2, lambda2=0.0), wq=w)
    assert rel(o3, o1) < 2e-3
    # (d) linearity in V (the softmax does not depend on V)
    kva, kvb = kv.clone(), kv.clone()
    kvb[..., C:] = rnd(B, Tk, C, dtype=BF16, seed=5)
    kvs = kv.clone()
    kvs[..., C:] = (0.5 * kva[..., C:].float() + 0.25 * kvb[..., C:].float()).to(BF16)
    oa = o1.float()
    ob = ops.xattn(h, ops.xattn_pack_kv(kvb, heads), wq=w).float()
    os_ = ops.xattn(h, ops.xattn_pack_kv(kvs, heads), wq=w).float()
    assert rel(os_, 0.5 * oa + 0.25 * ob) < 1e-2


def test_conv3x3_splitk_deterministic_and_linear():
    """Split-K convolution at the benchmark batch (16 x 8 x 8, 1280 -> 1280): two runs are bit-identical (fixed slice
    order, integer-atomic statistics) and the result is linear in the input."""
    B, H, C = 16, 8, 1280
    x1, x2 = rnd(B, H, H, C, dtype=BF16), rnd(B, H, H, C, dtype=BF16, seed=9)
    w = ops.pack_conv3x3(rnd(C, C, 3, 3, scale=(9 * C) ** -0.5, seed=1), BF16)
    st1 = torch.zeros(B * C * 2, device=DEV, dtype=torch.int64)
    st2 = torch.zeros_like(st1)
    y1 = ops.conv3x3(x1, w, None, stats=st1)
    y1b = ops.conv3x3(x1, w, None, stats=st2)
    assert torch.equal(y1, y1b) and torch.equal(st1, st2)
    y2 = ops.conv3x3(x2, w, None)
    xs = (0.5 * x1.float() - 0.25 * x2.float()).to(BF16)
    ys = ops.conv3x3(xs, w, None)
    assert rel(ys, 0.5 * y1.float() - 0.25 * y2.float()) < 1e-2
