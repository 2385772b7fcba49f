"""Thin torch-tensor front end over the C ABI (include/c2d.h).  PyTorch is used for device memory and
streams only; every function marshals raw device pointers into libc2d on the current CUDA stream.

Layout: activations are channels-last token tensors [B, N, C] (N = H*W); see c2d.h.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import lib, check

IMPL_AUTO, IMPL_SIMT, IMPL_TCGEN05 = _lib.IMPL_AUTO, _lib.IMPL_SIMT, _lib.IMPL_TCGEN05
ACT_NONE, ACT_GELU, ACT_SILU, ACT_RELU = _lib.ACT_NONE, _lib.ACT_GELU, _lib.ACT_SILU, _lib.ACT_RELU
AUDIO_ADD, AUDIO_CONCAT = _lib.AUDIO_ADD, _lib.AUDIO_CONCAT


def require_cuda(x, who: str) -> None:
    """CUDA-only guard of the drop-in modules: the library has no CPU path and says so instead of falling back."""
    if not x.is_cuda:
        raise _lib.C2DError(f"{who} runs on CUDA only (libc2d has no CPU path); got a {x.device} tensor")


# Per-op profiling for bench.py's roofline: when PROFILE is a list, every op appends
# (kernel_name, flops, bytes, start_event, end_event) recorded on the launching stream.
PROFILE = None


class _Timed:
    __slots__ = ("flops", "nbytes", "e0")

    def __init__(self, flops=0.0, nbytes=0.0):
        self.flops, self.nbytes = flops, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append(((lib.c2d_last_kernel() or b"?").decode(), float(self.flops), float(self.nbytes), self.e0, e1))
        return False


def _nb(*tensors) -> float:
    return float(sum(t.numel() * t.element_size() for t in tensors if t is not None))


_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"libc2d takes float32 or bfloat16 tensors, got {t.dtype}") from None


def _dev(t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise _lib.C2DError("libc2d has no CPU path: tensor is on " + str(t.device))
    _lib.ensure_init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous float32 tensor")
    return t


def _rows(x: torch.Tensor):
    """View [..., K] as (M, K, ld) without copying; requires unit inner stride and uniform row stride."""
    if x.stride(-1) != 1:
        raise ValueError("inner dimension must be contiguous")
    K = x.shape[-1]
    if x.dim() == 1:
        return 1, K, K
    M = 1
    for s in x.shape[:-1]:
        M *= s
    # row stride = stride of the innermost leading dim with extent > 1 (extent-1 dims may carry any stride,
    # e.g. numpy's 0 for a[None]); the remaining leading dims must collapse onto it
    ld, exp = K, None
    for dim in range(x.dim() - 2, -1, -1):
        if x.shape[dim] == 1:
            continue
        if exp is None:
            ld = x.stride(dim)
            exp = ld * x.shape[dim]
        else:
            if x.stride(dim) != exp:
                raise ValueError(f"tensor with shape {tuple(x.shape)} strides {x.stride()} is not a strided row matrix")
            exp *= x.shape[dim]
    return M, K, ld


# ------------------------------------------------------------------------------------------------
class LNFold:
    """LayerNorm folded into the next linear layer (pack_lnfold): gamma-scaled weight, its column sums, folded bias."""
    __slots__ = ("w", "colsum", "bias", "eps")

    def __init__(self, w, colsum, bias, eps):
        self.w, self.colsum, self.bias, self.eps = w, colsum, bias, float(eps)


def linear(x, w, bias=None, *, act=ACT_NONE, residual=None, rowvec=None, rows_per_vec=1, out=None,
           impl=IMPL_AUTO, x2=None, stats=None, stats_rows=0, row_stats=None, ln=None, ln_stats=None):
    """y = act([x | x2] @ w.T + bias + rowvec[row // rows_per_vec]) + residual.   w: [N, K] (nn.Linear layout).
    x2: optional second source concatenated along K (bf16 tcgen05 path).  stats: optional int64 [B, N, 2] channel
    statistics accumulator (rows_per_image = stats_rows) filled by the epilogue for group_norm_apply.
    row_stats: optional int64 [M, 2] per-row statistics accumulator of y (for a following folded LayerNorm).
    ln / ln_stats: LNFold of the preceding LayerNorm and the int64 [M, 2] row statistics of x: computes
    LayerNorm(x) @ W.T + b without a LayerNorm kernel (w / bias arguments are taken from `ln`)."""
    if ln is not None:
        w, bias = ln.w, ln.bias
    _dev(x)
    M, K1, ldx = _rows(x)
    K, ldx2 = K1, 0
    if x2 is not None:
        M2, K2, ldx2 = _rows(x2)
        assert M2 == M and x2.dtype == x.dtype
        K = K1 + K2
    N = w.shape[0]
    assert w.shape[1] == K and w.is_contiguous() and w.dtype == x.dtype, (w.shape, K, w.dtype, x.dtype)
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=x.dtype)
    Mo, No, ldy = _rows(out)
    assert Mo == M and No == N and out.dtype == x.dtype
    ldr = 0
    if residual is not None:
        Mr, Nr, ldr = _rows(residual)
        assert Mr == M and Nr == N and residual.dtype == x.dtype
    with _Timed(2.0 * M * N * K, _nb(w, residual) + (M * K + M * N) * x.element_size()):
        if x2 is None and stats is None and row_stats is None and ln is None:
            check(lib.c2d_linear(x.data_ptr(), w.data_ptr(), _ptr(_f32(bias, "bias")), _ptr(_f32(rowvec, "rowvec")),
                                 int(rows_per_vec), _ptr(residual), out.data_ptr(), M, N, K, ldx, ldy, ldr, act, _dt(x),
                                 impl, _stream()), "linear")
        else:
            if stats is not None:
                assert stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() == (M // stats_rows) * N * 2
            if row_stats is not None:
                assert row_stats.dtype == torch.int64 and row_stats.is_contiguous() and row_stats.numel() == M * 2
            if ln is not None:
                assert ln_stats is not None and ln_stats.dtype == torch.int64 and ln_stats.numel() == M * 2
            check(lib.c2d_linear_ex(x.data_ptr(), _ptr(x2), K1, ldx2, w.data_ptr(), _ptr(_f32(bias, "bias")),
                                    _ptr(_f32(rowvec, "rowvec")), int(rows_per_vec), _ptr(residual), out.data_ptr(), M, N, K,
                                    ldx, ldy, ldr, act, _ptr(stats), int(stats_rows), _ptr(row_stats),
                                    _ptr(ln_stats) if ln is not None else None,
                                    _ptr(_f32(ln.colsum, "colsum")) if ln is not None else None,
                                    ln.eps if ln is not None else 0.0, _dt(x), _stream()), "linear_ex")
    return out


def geglu_linear(x, w_packed, bias_packed, *, out=None, impl=IMPL_AUTO, ln=None, ln_stats=None):
    """Fused GEGLU projection; w_packed/bias_packed from pack_geglu().  ln / ln_stats: folded LayerNorm (packed LNFold)."""
    _dev(x)
    if ln is not None:
        w_packed, bias_packed = ln.w, ln.bias
    M, K, ldx = _rows(x)
    assert ldx == K, "geglu_linear takes a dense x"
    F = w_packed.shape[0] // 2
    if out is None:
        out = torch.empty(*x.shape[:-1], F, device=x.device, dtype=x.dtype)
    with _Timed(4.0 * M * F * K, _nb(x, w_packed, out)):
        if ln is None:
            check(lib.c2d_geglu_linear(x.data_ptr(), w_packed.data_ptr(), _ptr(_f32(bias_packed, "bias")), out.data_ptr(),
                                       M, F, K, 1, _dt(x), impl, _stream()), "geglu_linear")
        else:
            assert ln_stats is not None and ln_stats.dtype == torch.int64 and ln_stats.numel() == M * 2
            check(lib.c2d_geglu_linear_ex(x.data_ptr(), w_packed.data_ptr(), _ptr(_f32(bias_packed, "bias")),
                                          ln_stats.data_ptr(), _f32(ln.colsum, "colsum").data_ptr(), ln.eps, out.data_ptr(),
                                          M, F, K, _dt(x), _stream()), "geglu_linear_ex")
    return out


def conv3x3(x, w_packed, bias=None, *, rowvec=None, residual=None, stride=1, upsample=False, out=None,
            impl=IMPL_AUTO, stats=None):
    """x [B,H,W,Cin] NHWC; w_packed [Cout,3,3,Cin]; returns [B,Ho,Wo,Cout].  stats: optional int64 [B,Cout,2]
    channel-statistics accumulator filled by the epilogue (bf16 tcgen05 path)."""
    _dev(x)
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    assert x.is_contiguous() and w_packed.is_contiguous() and tuple(w_packed.shape[1:]) == (3, 3, Cin)
    Hin, Win = (2 * H, 2 * W) if upsample else (H, W)
    Ho, Wo = (Hin - 1) // stride + 1, (Win - 1) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, Cout, device=x.device, dtype=x.dtype)
    assert out.is_contiguous() and out.numel() == B * Ho * Wo * Cout
    if residual is not None:
        assert residual.is_contiguous() and residual.numel() == out.numel() and residual.dtype == x.dtype
    with _Timed(2.0 * B * Ho * Wo * Cout * 9 * Cin, _nb(x, w_packed, out, residual)):
        if stats is None:
            check(lib.c2d_conv3x3(x.data_ptr(), w_packed.data_ptr(), _ptr(_f32(bias, "bias")), _ptr(_f32(rowvec, "rowvec")),
                                  _ptr(residual), out.data_ptr(), B, H, W, Cin, Cout, stride, int(bool(upsample)), _dt(x),
                                  impl, _stream()), "conv3x3")
        else:
            assert not upsample and stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() == B * Cout * 2
            check(lib.c2d_conv3x3_ex(x.data_ptr(), w_packed.data_ptr(), _ptr(_f32(bias, "bias")),
                                     _ptr(_f32(rowvec, "rowvec")), _ptr(residual), out.data_ptr(), B, H, W, Cin, Cout,
                                     stride, stats.data_ptr(), _dt(x), _stream()), "conv3x3_ex")
    return out


def conv3x3_down(x, w_packed, bias=None, *, out=None, impl=IMPL_AUTO, stats=None):
    """diffusers Downsample2D(padding=0): zero-pad right / bottom by one, 3x3 convolution with stride 2.
    x [B,H,W,Cin] NHWC (H, W even) -> [B,H/2,W/2,Cout]."""
    _dev(x)
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    assert x.is_contiguous() and w_packed.is_contiguous() and tuple(w_packed.shape[1:]) == (3, 3, Cin) and H % 2 == 0 and W % 2 == 0
    if out is None:
        out = torch.empty(B, H // 2, W // 2, Cout, device=x.device, dtype=x.dtype)
    if stats is not None:
        assert stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() == B * Cout * 2
    with _Timed(2.0 * B * (H // 2) * (W // 2) * Cout * 9 * Cin, _nb(x, w_packed, out)):
        check(lib.c2d_conv3x3_down(x.data_ptr(), w_packed.data_ptr(), _ptr(_f32(bias, "bias")), out.data_ptr(), B, H, W, Cin, Cout,
                                   _ptr(stats), _dt(x), impl, _stream()), "conv3x3_down")
    return out


_gn_ws = {}


def _gn_workspace(device, n_doubles: int) -> torch.Tensor:
    """fp32-mode GroupNorm statistics scratch, one per (device, stream) so that concurrent streams never share it."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _gn_ws.get(key)
    if ws is None or ws.numel() < n_doubles:
        ws = torch.empty(max(n_doubles, 8192), device=device, dtype=torch.float64)
        _gn_ws[key] = ws
    return ws


def group_norm(x, gamma, beta, groups=32, eps=1e-5, silu=False, *, x2=None, raw_cat=None, out=None):
    """x [B,N,C1] (+ optional x2 [B,N,C2] concatenated on channels) -> [B,N,C1+C2]."""
    _dev(x)
    B, C1 = x.shape[0], x.shape[-1]
    N = x.numel() // (B * C1)
    C2 = 0 if x2 is None else x2.shape[-1]
    assert x.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.dtype == x.dtype))
    if out is None:
        out = torch.empty(*x.shape[:-1], C1 + C2, device=x.device, dtype=x.dtype)
    ws = _gn_workspace(x.device, B * groups * 2)
    with _Timed(0.0, 2 * _nb(x, x2) + _nb(out, raw_cat)):     # stats pass + apply pass read the input twice
        check(lib.c2d_group_norm(x.data_ptr(), _ptr(x2), _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(),
                                 out.data_ptr(), _ptr(raw_cat), ws.data_ptr(), B, N, C1, C2, groups, float(eps),
                                 int(bool(silu)), _dt(x), _stream()), "group_norm")
    return out


def channel_stats(x, stats):
    """Accumulate per-channel (sum, sumsq) of x [B,N,C] into the int64 [B,C,2] fixed-point accumulator `stats`."""
    _dev(x)
    B, C = x.shape[0], x.shape[-1]
    N = x.numel() // (B * C)
    assert x.is_contiguous() and stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() == B * C * 2
    with _Timed(0.0, _nb(x)):
        check(lib.c2d_channel_stats(x.data_ptr(), stats.data_ptr(), B, N, C, _dt(x), _stream()), "channel_stats")
    return stats


def group_norm_apply(x, stats, gamma, beta, groups=32, eps=1e-5, silu=False, *, x2=None, stats2=None, out=None):
    """GroupNorm(+SiLU) of cat([x, x2], -1) from producer-side channel statistics: one pass, no statistics read."""
    _dev(x)
    B, C1 = x.shape[0], x.shape[-1]
    N = x.numel() // (B * C1)
    C2 = 0 if x2 is None else x2.shape[-1]
    assert x.is_contiguous() and (x2 is None or (x2.is_contiguous() and x2.dtype == x.dtype and stats2 is not None))
    assert stats.dtype == torch.int64 and stats.numel() == B * C1 * 2
    assert stats2 is None or (stats2.dtype == torch.int64 and stats2.numel() == B * C2 * 2)
    if out is None:
        out = torch.empty(*x.shape[:-1], C1 + C2, device=x.device, dtype=x.dtype)
    with _Timed(0.0, _nb(x, x2, out)):
        check(lib.c2d_group_norm_apply(x.data_ptr(), _ptr(x2), stats.data_ptr(), _ptr(stats2),
                                       _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(), out.data_ptr(),
                                       B, N, C1, C2, groups, float(eps), int(bool(silu)), _dt(x), _stream()),
              "group_norm_apply")
    return out


def layer_norm(x, gamma, beta, eps=1e-5, *, out=None):
    _dev(x)
    assert x.is_contiguous()
    C = x.shape[-1]
    M = x.numel() // C
    if out is None:
        out = torch.empty_like(x)
    with _Timed(0.0, _nb(x, out)):
        check(lib.c2d_layer_norm(x.data_ptr(), _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(),
                                 out.data_ptr(), M, C, float(eps), _dt(x), _stream()), "layer_norm")
    return out


def attention(q, k, v, heads: int, *, scale: Optional[float] = None, mask=None, out=None, impl=IMPL_AUTO, lse=None):
    """softmax(q k^T * scale) v per head.  q [B,Nq,heads*d], k/v [B,Nkv,heads*d]; strided views (e.g. slices
    of a packed QKV buffer) are accepted as long as the inner dim is contiguous.
    lse (training forward): fp32 [B,heads,Nq] buffer for the rows' log-sum-exp; the call then returns (out, written) where
    `written` tells whether the kernel that ran produces it (attention_bwd(..., lse=...) skips a sweep when it did)."""
    _dev(q)
    B, Nq, C = q.shape
    Nkv = k.shape[1]
    d = C // heads
    assert k.shape[2] == C and v.shape[2] == C and v.shape[1] == Nkv
    assert q.stride(2) == 1 and k.stride(2) == 1 and v.stride(2) == 1
    if out is None:
        out = torch.empty(B, Nq, C, device=q.device, dtype=q.dtype)
    assert out.stride(2) == 1
    if scale is None:
        scale = d ** -0.5
    if mask is not None:
        assert mask.dtype in (torch.bool, torch.uint8) and mask.is_contiguous() and tuple(mask.shape) == (B, Nkv)
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and tuple(lse.shape) == (B, heads, Nq)
        written = ctypes.c_int(0)
        with _Timed(4.0 * B * heads * Nq * Nkv * d, (2 * B * Nq * C + 2 * B * Nkv * C) * q.element_size()):
            check(lib.c2d_attention_lse(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, heads, Nq, Nkv, d,
                                        q.stride(1), k.stride(1), v.stride(1), out.stride(1), q.stride(0), k.stride(0),
                                        v.stride(0), out.stride(0), float(scale), _ptr(mask), lse.data_ptr(), ctypes.byref(written),
                                        _dt(q), impl, _stream()), "attention")
        return out, bool(written.value)
    with _Timed(4.0 * B * heads * Nq * Nkv * d, (2 * B * Nq * C + 2 * B * Nkv * C) * q.element_size()):
        check(lib.c2d_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, heads, Nq, Nkv, d,
                                q.stride(1), k.stride(1), v.stride(1), out.stride(1), q.stride(0), k.stride(0),
                                v.stride(0), out.stride(0), float(scale), _ptr(mask), _dt(q), impl, _stream()),
              "attention")
    return out


class XattnKV:
    """Packed, step-invariant K / V cache of one cross-attention site (c2d_xattn_pack_kv): per (batch, head) the exact
    shared-memory image the fused kernel fetches with one bulk copy.  T text(+audio) keys, T2 decoupled audio keys."""
    __slots__ = ("packed", "B", "C", "heads", "T", "T2", "kv", "lambda2")

    def __init__(self, packed, B, C, heads, T, T2, kv=None, lambda2=1.0):
        self.packed, self.B, self.C, self.heads, self.T, self.T2 = packed, B, C, heads, T, T2
        self.lambda2 = float(lambda2)      # scale of the decoupled second branch
        self.kv = kv          # the unpacked [B,T,2C] tensor, kept for token counts the fused kernel does not take

    @property
    def shape(self):          # identifies the cache layout (the sampler keys its captured graphs on it)
        # lambda2 is a kernel ARGUMENT (baked into a captured graph), so it is part of the identity
        return ("xattn_kv", self.B, self.C, self.heads, self.T, self.T2, self.lambda2)

    def copy_(self, other: "XattnKV"):
        """Refresh the buffers a CUDA graph was captured on with a new image batch's cache."""
        assert self.shape == other.shape
        if self.packed is not None:
            self.packed.copy_(other.packed)
        if self.kv is not None and other.kv is not None:
            self.kv.copy_(other.kv)
        return self


def xattn_supported(x, heads: int, T: int, T2: int = 0) -> bool:
    """True when the fused cross-attention kernel takes this site (bf16, SD-1.5 shapes); no device work."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 3:
        return False
    return bool(lib.c2d_xattn_supported(x.shape[-1], int(heads), x.shape[1], int(T), int(T2), _lib.BF16))


def xattn_packable(C: int, heads: int, T: int, dtype, T2: int = 0) -> bool:
    """True when c2d_xattn_pack_kv takes this site (bf16, head dims / key counts inside the fused kernel)."""
    if dtype != torch.bfloat16:
        return False
    return int(lib.c2d_xattn_packed_bytes(int(C), int(heads), int(T), int(T2))) > 0


def xattn_pack_kv(kv, heads: int, kv2=None, lambda2: float = 1.0) -> XattnKV:
    """kv [B,T,2C] = [K | V] of a site (text keys with the audio injected); kv2 [B,T2,2C]: decoupled second branch,
    scaled by lambda2 when it is added."""
    _dev(kv)
    B, T, C2 = kv.shape
    C = C2 // 2
    assert kv.stride(2) == 1 and kv.dtype == torch.bfloat16, "the fused cross-attention kernel is bf16 only"
    T2 = 0
    if kv2 is not None:
        T2 = kv2.shape[1]
        assert kv2.shape[0] == B and kv2.shape[2] == C2 and kv2.stride(2) == 1 and kv2.dtype == kv.dtype
    nbytes = int(lib.c2d_xattn_packed_bytes(C, int(heads), T, T2))
    if nbytes <= 0:
        raise _lib.C2DError(f"xattn_pack_kv: shape outside the fused kernel (C={C} heads={heads} T={T} T2={T2})")
    packed = torch.empty(B, nbytes // 2, device=kv.device, dtype=kv.dtype)
    check(lib.c2d_xattn_pack_kv(kv.data_ptr(), kv[..., C:].data_ptr(), kv.stride(1), kv.stride(0), T,
                                None if kv2 is None else kv2.data_ptr(), None if kv2 is None else kv2[..., C:].data_ptr(),
                                0 if kv2 is None else kv2.stride(1), 0 if kv2 is None else kv2.stride(0), T2,
                                packed.data_ptr(), B, C, int(heads), _dt(kv), _stream()), "xattn_pack_kv")
    return XattnKV(packed, B, C, int(heads), T, T2, kv if kv2 is None else None, lambda2)


def xattn(x, kvp: XattnKV, *, wq=None, q_bias=None, ln=None, ln_stats=None, scale: Optional[float] = None,
          lambda2: Optional[float] = None, out=None):
    """Fused cross-attention site, per-step part (c2d_xattn_fwd): to_q (+ folded LayerNorm) + softmax(q k^T) v in one
    kernel.  x [B,Nq,C]; kvp from xattn_pack_kv; ln / ln_stats: LNFold of the preceding LayerNorm and the int64
    [B*Nq,2] row statistics of x (then wq / q_bias come from `ln`); lambda2 scales the decoupled second branch
    (default: the value the cache was packed with).
    Returns o [B,Nq,C] (input of to_out)."""
    _dev(x)
    if ln is not None:
        wq, q_bias = ln.w, ln.bias
        assert ln_stats is not None and ln_stats.dtype == torch.int64 and ln_stats.numel() == x.shape[0] * x.shape[1] * 2
    B, Nq, C = x.shape
    assert kvp.B == B and kvp.C == C and kvp.packed.dtype == x.dtype
    assert x.stride(2) == 1 and x.stride(0) == Nq * x.stride(1), "x must be a uniformly strided [B*Nq, C] row matrix"
    assert wq.shape == (C, C) and wq.is_contiguous() and wq.dtype == x.dtype
    d = C // kvp.heads
    if scale is None:
        scale = d ** -0.5
    if lambda2 is None:
        lambda2 = kvp.lambda2
    if out is None:
        out = torch.empty(B, Nq, C, device=x.device, dtype=x.dtype)
    assert out.stride(2) == 1 and out.stride(0) == Nq * out.stride(1)
    flops = 2.0 * B * Nq * C * C + 4.0 * B * Nq * (kvp.T + kvp.T2) * C
    with _Timed(flops, _nb(x, wq, kvp.packed, out)):
        check(lib.c2d_xattn_fwd(x.data_ptr(), x.stride(1), wq.data_ptr(), _ptr(_f32(q_bias, "q_bias")),
                                _ptr(ln_stats) if ln is not None else None,
                                _ptr(_f32(ln.colsum, "colsum")) if ln is not None else None,
                                ln.eps if ln is not None else 0.0, kvp.packed.data_ptr(), kvp.T, kvp.T2, float(lambda2),
                                out.data_ptr(), out.stride(1), B, Nq, C, kvp.heads, float(scale), _dt(x), _stream()),
              "xattn_fwd")
    return out


def audio_context(ehs, audio, w1, b1, w2, b2, alpha, mode: int, *, out=None):
    """AudioAttnProcessor context step (add / concat); see c2d.h."""
    _dev(ehs)
    B, T, D = ehs.shape
    K, Da = audio.shape[1], audio.shape[2]
    Hb = w1.shape[0]
    assert ehs.is_contiguous() and audio.is_contiguous() and audio.dtype == ehs.dtype
    assert w1.dtype == ehs.dtype and w2.dtype == ehs.dtype and w1.is_contiguous() and w2.is_contiguous()
    Tout = T if mode == AUDIO_ADD else T + min(K, 4)
    if out is None:
        out = torch.empty(B, Tout, D, device=ehs.device, dtype=ehs.dtype)
    check(lib.c2d_audio_context(ehs.data_ptr(), audio.data_ptr(), w1.data_ptr(), _f32(b1, "b1").data_ptr(),
                                w2.data_ptr(), _f32(b2, "b2").data_ptr(), _ptr(_f32(alpha, "alpha")), out.data_ptr(),
                                B, T, D, K, Da, Hb, mode, _dt(ehs), _stream()), "audio_context")
    return out


# ------------------------------------------------------------------------------------------------
def timestep_embedding(t, dim: int = 320, *, out=None):
    _dev(t)
    t = _f32(t, "t")
    B = t.numel()
    if out is None:
        out = torch.empty(B, dim, device=t.device, dtype=torch.float32)
    check(lib.c2d_timestep_embedding(t.data_ptr(), out.data_ptr(), B, dim, _stream()), "timestep_embedding")
    return out


def unary(x, act=ACT_NONE, *, out_dtype=None, out=None):
    """Elementwise activation and/or dtype cast."""
    _dev(x)
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    check(lib.c2d_unary(x.data_ptr(), out.data_ptr(), x.numel(), act, _dt(x), _dt(out), _stream()), "unary")
    return out


def cast(x, dtype, *, out=None):
    return unary(x, ACT_NONE, out_dtype=dtype, out=out)


def add(a, b, *, out=None):
    _dev(a)
    assert a.is_contiguous() and b.is_contiguous() and a.shape == b.shape and a.dtype == b.dtype
    if out is None:
        out = torch.empty_like(a)
    check(lib.c2d_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _dt(a), _stream()), "add")
    return out


def geglu(x, *, out=None):
    _dev(x)
    assert x.is_contiguous()
    F = x.shape[-1] // 2
    M = x.numel() // (2 * F)
    if out is None:
        out = torch.empty(*x.shape[:-1], F, device=x.device, dtype=x.dtype)
    check(lib.c2d_geglu(x.data_ptr(), out.data_ptr(), M, F, _dt(x), _stream()), "geglu")
    return out


def upsample2x(x, *, out=None):
    _dev(x)
    B, H, W, Cc = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(B, 2 * H, 2 * W, Cc, device=x.device, dtype=x.dtype)
    with _Timed(0.0, _nb(x, out)):
        check(lib.c2d_upsample2x(x.data_ptr(), out.data_ptr(), B, H, W, Cc, _dt(x), _stream()), "upsample2x")
    return out


def concat(x1, x2, *, out=None):
    _dev(x1)
    assert x1.is_contiguous() and x2.is_contiguous() and x1.shape[:-1] == x2.shape[:-1] and x1.dtype == x2.dtype
    C1, C2 = x1.shape[-1], x2.shape[-1]
    rows = x1.numel() // C1
    if out is None:
        out = torch.empty(*x1.shape[:-1], C1 + C2, device=x1.device, dtype=x1.dtype)
    check(lib.c2d_concat(x1.data_ptr(), x2.data_ptr(), out.data_ptr(), rows, C1, C2, _dt(x1), _stream()), "concat")
    return out


def nchw_to_nhwc(x, dtype, *, out=None):
    """fp32 [B,C,H,W] -> dtype [B,H,W,C]."""
    _dev(x)
    x = _f32(x, "x")
    B, Cc, H, W = x.shape
    if out is None:
        out = torch.empty(B, H, W, Cc, device=x.device, dtype=dtype)
    check(lib.c2d_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), B, Cc, H * W, _dt(out), _stream()), "nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, *, out=None):
    """dtype [B,H,W,C] -> fp32 [B,C,H,W]."""
    _dev(x)
    B, H, W, Cc = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(B, Cc, H, W, device=x.device, dtype=torch.float32)
    check(lib.c2d_nhwc_to_nchw(x.data_ptr(), _f32(out, "out").data_ptr(), B, Cc, H * W, _dt(x), _stream()),
          "nhwc_to_nchw")
    return out


def cfg_sched_step(eps2, x, xin2, guidance: float, coef, trace=None):
    """In place: x <- ca*x + cb*(eps_u + g (eps_c - eps_u)); xin2 <- cat[x, x] * in_scale (NHWC, eps2.dtype).
    coef: device fp32 tensor [3] = (ca, cb, in_scale).  xin2 may have 8 channels (4 real + 4 zero padding the caller
    initialised): only the real ones are written."""
    _dev(eps2)
    B = x.shape[0]
    HW = x.shape[2] * x.shape[3]
    assert eps2.is_contiguous() and xin2.is_contiguous() and eps2.shape[0] == 2 * B and xin2.dtype == eps2.dtype
    assert coef.is_cuda and coef.numel() >= 3 and eps2.shape[-1] == 4 and xin2.shape[-1] in (4, 8)
    check(lib.c2d_cfg_sched_step(eps2.data_ptr(), _f32(x, "x").data_ptr(), xin2.data_ptr(), _ptr(_f32(trace, "trace")),
                                 B, HW, float(guidance), _f32(coef, "coef").data_ptr(), int(xin2.shape[-1]), _dt(eps2),
                                 _stream()), "cfg_sched_step")
    return x


def softmax_rows(x, scale: float = 1.0, *, out=None):
    _dev(x)
    assert x.is_contiguous()
    N = x.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    check(lib.c2d_softmax_rows(x.data_ptr(), out.data_ptr(), x.numel() // N, N, float(scale), _dt(x), _stream()),
          "softmax_rows")
    return out


def transpose(x, *, out=None):
    """[b, R, C] -> [b, C, R] (materialised)."""
    _dev(x)
    assert x.is_contiguous() and x.dim() == 3
    b, R, Cc = x.shape
    if out is None:
        out = torch.empty(b, Cc, R, device=x.device, dtype=x.dtype)
    check(lib.c2d_transpose(x.data_ptr(), out.data_ptr(), b, R, Cc, _dt(x), _stream()), "transpose")
    return out


def pack_conv3x3(w, dtype):
    """[Cout,Cin,3,3] fp32 (diffusers) -> [Cout,3,3,Cin] dtype."""
    _dev(w)
    w = _f32(w, "w")
    Cout, Cin = w.shape[0], w.shape[1]
    out = torch.empty(Cout, 3, 3, Cin, device=w.device, dtype=dtype)
    check(lib.c2d_pack_conv3x3(w.data_ptr(), out.data_ptr(), Cout, Cin, _dt(out), _stream()), "pack_conv3x3")
    return out


def pack_lnfold(w, gamma, beta, bias, dtype, eps=1e-5, out_dtype=None):
    """Fold LayerNorm(gamma, beta) into the linear layer (w [N,K] fp32, bias [N] fp32 or None) that follows it.
    Returns an LNFold whose weight has dtype `out_dtype` (default `dtype`); column sums are taken over the
    `dtype`-rounded (bf16) weights."""
    _dev(w)
    w = _f32(w, "w")
    N, K = w.shape
    wo = torch.empty(N, K, device=w.device, dtype=out_dtype or dtype)
    cs = torch.empty(N, device=w.device, dtype=torch.float32)
    bo = torch.empty(N, device=w.device, dtype=torch.float32)
    check(lib.c2d_pack_lnfold(w.data_ptr(), _f32(gamma, "gamma").data_ptr(), _f32(beta, "beta").data_ptr(),
                              _ptr(_f32(bias, "bias")), wo.data_ptr(), cs.data_ptr(), bo.data_ptr(), N, K, _dt(wo),
                              _stream()), "pack_lnfold")
    return LNFold(wo, cs, bo, eps)


def pack_geglu(w, bias, dtype):
    """diffusers ff.net.0.proj [2F,K] fp32 (+bias [2F]) -> row-interleaved (64-row a/g blocks) dtype weights."""
    _dev(w)
    w = _f32(w, "w")
    F, K = w.shape[0] // 2, w.shape[1]
    wo = torch.empty(2 * F, K, device=w.device, dtype=dtype)
    bo = torch.empty(2 * F, device=w.device, dtype=torch.float32) if bias is not None else None
    check(lib.c2d_pack_geglu(w.data_ptr(), _ptr(_f32(bias, "bias")), wo.data_ptr(), _ptr(bo), F, K, _dt(wo),
                             _stream()), "pack_geglu")
    return wo, bo


# ------------------------------------------------------------------------------------------------
# audio-conditioning side
# ------------------------------------------------------------------------------------------------
def bcast_add(a, b, B: int, K: int, D: int, a_mode: int, b_mode: int, *, out=None):
    """out[B,K,D] = a + b; mode 0 = [B,K,D], 1 = [B,1,D] (broadcast over tokens), 2 = [1,K,D] (over batch)."""
    _dev(a)
    assert a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype
    if out is None:
        out = torch.empty(B, K, D, device=a.device, dtype=a.dtype)
    check(lib.c2d_bcast_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), B, K, D, a_mode, b_mode, _dt(a), _stream()),
          "bcast_add")
    return out


def hier_assign(tokens, anchors, w1, b1, w2, b2, temperature):
    """Soft level assignments fp32 [B,K,L]; tokens [B,K,D]; temperature: device fp32 scalar tensor."""
    _dev(tokens)
    B, K, D = tokens.shape
    L, Hg = anchors.shape[0], w1.shape[0]
    assert tokens.is_contiguous() and anchors.is_contiguous() and w1.is_contiguous() and w2.is_contiguous()
    out = torch.empty(B, K, L, device=tokens.device, dtype=torch.float32)
    check(lib.c2d_hier_assign(tokens.data_ptr(), anchors.data_ptr(), w1.data_ptr(), _f32(b1, "b1").data_ptr(),
                              w2.data_ptr(), _f32(b2, "b2").data_ptr(), _f32(temperature, "temperature").data_ptr(),
                              out.data_ptr(), B * K, D, L, Hg, _dt(tokens), _stream()), "hier_assign")
    return out


def hier_route(tok10, assign, hw, routing, gates):
    """Returns (early, mid, late) routed tokens, each [B,K,D] in tok10.dtype."""
    _dev(tok10)
    B, K, D = tok10.shape
    assert tok10.is_contiguous() and assign.shape[-1] == 3
    outs = [torch.empty_like(tok10) for _ in range(3)]
    check(lib.c2d_hier_route(tok10.data_ptr(), _f32(assign, "assign").data_ptr(), _ptr(_f32(hw, "hw")),
                             _f32(routing, "routing").data_ptr(), _f32(gates, "gates").data_ptr(),
                             outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), B, K, D, _dt(tok10),
                             _stream()), "hier_route")
    return tuple(outs)


def norm_scale(x, target: float = 60.0, per_sample: bool = False, *, out=None):
    _dev(x)
    B, K, D = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(lib.c2d_norm_scale(x.data_ptr(), out.data_ptr(), B, K, D, float(target), int(bool(per_sample)), _dt(x),
                             _stream()), "norm_scale")
    return out


def legacy_combine(fg, bg, amb, hierarchy_weights, D: int):
    """cat(fg*w0, bg*w1, amb*w2) with w = softmax(hierarchy_weights); fg [B,nf*D] etc. -> [B,nf+nb+na,D]."""
    _dev(fg)
    B = fg.shape[0]
    nf, nb, na = fg.shape[1] // D, bg.shape[1] // D, amb.shape[1] // D
    assert fg.is_contiguous() and bg.is_contiguous() and amb.is_contiguous()
    out = torch.empty(B, nf + nb + na, D, device=fg.device, dtype=fg.dtype)
    check(lib.c2d_legacy_combine(fg.data_ptr(), bg.data_ptr(), amb.data_ptr(),
                                 _f32(hierarchy_weights, "hierarchy_weights").data_ptr(), out.data_ptr(), B, nf, nb,
                                 na, D, _dt(fg), _stream()), "legacy_combine")
    return out


# ------------------------------------------------------------------------------------------------
# CLAP HTSAT audio tower (log-mel front end, Swin window attention, patch merging, pooling)
# ------------------------------------------------------------------------------------------------
def stft_frames(wave, window, hop: int, n_frames: int, *, out=None):
    """wave fp32 [B,T], window fp32 [n_fft] -> frames fp32 [B*n_frames, n_fft] (centered, reflect-padded, windowed)."""
    _dev(wave)
    B, T = wave.shape
    n_fft = window.numel()
    assert wave.is_contiguous() and window.is_contiguous()
    if out is None:
        out = torch.empty(B * n_frames, n_fft, device=wave.device, dtype=torch.float32)
    check(lib.c2d_stft_frames(_f32(wave, "wave").data_ptr(), _f32(window, "window").data_ptr(), out.data_ptr(), B, T, n_fft,
                              int(hop), int(n_frames), _stream()), "stft_frames")
    return out


def stft_frames_split(wave, window, hop: int, n_frames: int, *, out=None):
    """wave fp32 [B,T], window fp32 [n_fft] -> bf16 [B*n_frames, 3*n_fft] = [hi | lo | hi] (split-bf16 operand of the
    tensor-core DFT: meets constant-matrix rows [HI | HI | LO])."""
    _dev(wave)
    B, T = wave.shape
    n_fft = window.numel()
    assert wave.is_contiguous() and window.is_contiguous()
    if out is None:
        out = torch.empty(B * n_frames, 3 * n_fft, device=wave.device, dtype=torch.bfloat16)
    check(lib.c2d_stft_frames_split(_f32(wave, "wave").data_ptr(), _f32(window, "window").data_ptr(), out.data_ptr(), B, T, n_fft,
                                    int(hop), int(n_frames), _stream()), "stft_frames_split")
    return out


def power_spectrum(dft, *, nb: Optional[int] = None, im_off: Optional[int] = None, out=None):
    """dft [M, ld] (fp32 or bf16) with re(0..nb) at column 0 and im(0..nb) at column im_off -> fp32 [M, nb] = re^2 + im^2.
    Default layout: ld = 2*nb, im_off = nb."""
    _dev(dft)
    M, ld = dft.shape
    if nb is None:
        assert ld % 2 == 0
        nb = ld // 2
    if im_off is None:
        im_off = nb
    assert dft.is_contiguous() and ld >= im_off + nb
    if out is None:
        out = torch.empty(M, nb, device=dft.device, dtype=torch.float32)
    check(lib.c2d_power_spectrum(dft.data_ptr(), out.data_ptr(), M, int(nb), int(ld), int(im_off), _dt(dft), _stream()),
          "power_spectrum")
    return out


def log_mel_affine(x, a, b, floor: float = 1e-10, *, out=None):
    """10 log10(max(x, floor)) * a[f] + b[f] over x fp32 [M, F]."""
    _dev(x)
    M, Fm = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(lib.c2d_log_mel_affine(_f32(x, "x").data_ptr(), _f32(a, "a").data_ptr(), _f32(b, "b").data_ptr(), out.data_ptr(),
                                 M, Fm, float(floor), _stream()), "log_mel_affine")
    return out


def clap_patches(mel, dtype, *, out=None):
    """mel fp32 [B, n_frames, 64] (BatchNorm applied) -> [B*4096, 16] patch vectors in `dtype`."""
    _dev(mel)
    B, n_frames, n_mel = mel.shape
    assert mel.is_contiguous()
    if out is None:
        out = torch.empty(B * 4096, 16, device=mel.device, dtype=dtype)
    check(lib.c2d_clap_patches(_f32(mel, "mel").data_ptr(), out.data_ptr(), B, n_frames, n_mel, _dt(out), _stream()),
          "clap_patches")
    return out


def window_attention(qkv, bias, H: int, W: int, heads: int, shift: int, *, out=None):
    """Swin 8x8 window attention: qkv [B, H*W, 3C], bias fp32 [heads, 64, 64] -> [B, H*W, C]."""
    _dev(qkv)
    B, N, C3 = qkv.shape
    C = C3 // 3
    assert qkv.is_contiguous() and N == H * W and bias.is_contiguous() and tuple(bias.shape) == (heads, 64, 64)
    if out is None:
        out = torch.empty(B, N, C, device=qkv.device, dtype=qkv.dtype)
    d = C // heads
    with _Timed(4.0 * B * N * 64 * C, _nb(qkv, out)):
        check(lib.c2d_window_attention(qkv.data_ptr(), _f32(bias, "bias").data_ptr(), out.data_ptr(), B, H, W, C, heads,
                                       int(shift), float(d ** -0.5), _dt(qkv), _stream()), "window_attention")
    return out


def patch_merge(x, H: int, W: int, *, out=None):
    """x [B, H*W, C] -> [B, (H/2)*(W/2), 4C] (Swin patch-merging gather)."""
    _dev(x)
    B, N, C = x.shape
    assert x.is_contiguous() and N == H * W
    if out is None:
        out = torch.empty(B, (H // 2) * (W // 2), 4 * C, device=x.device, dtype=x.dtype)
    check(lib.c2d_patch_merge(x.data_ptr(), out.data_ptr(), B, H, W, C, _dt(x), _stream()), "patch_merge")
    return out


def token_mean(x, *, out=None):
    """x [B, N, C] -> fp32 [B, C]."""
    _dev(x)
    B, N, C = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(B, C, device=x.device, dtype=torch.float32)
    check(lib.c2d_token_mean(x.data_ptr(), out.data_ptr(), B, N, C, _dt(x), _stream()), "token_mean")
    return out


def l2_normalize(x, eps: float = 1e-12, *, out=None):
    """x fp32 [B, D] -> x / max(||x||, eps)."""
    _dev(x)
    B, D = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(lib.c2d_l2_normalize(_f32(x, "x").data_ptr(), out.data_ptr(), B, D, float(eps), _stream()), "l2_normalize")
    return out


# ======================================================================================================
# Backward / optimiser entry points of the stage-3 fine-tune step (include/c2d.h "stage-3 fine-tune step")
# ======================================================================================================
def group_norm_bwd(x, dy, gamma, beta, groups=32, eps=1e-5, silu=False, *, add=None, out=None, stats=None):
    """Adjoint of GroupNorm(+SiLU).  stats: the int64 [B,C,2] channel statistics of x from the forward pass
    (channel_stats / a GEMM epilogue); the bf16 path then skips its own statistics pass."""
    _dev(x)
    B, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (B * C)
    assert x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype
    assert add is None or (add.is_contiguous() and add.shape == x.shape and add.dtype == x.dtype)
    if out is None:
        out = torch.empty_like(x)
    # scratch of the three-pass bf16 path: int64 channel statistics + double adjoint sums, zeroed per call
    fast = x.dtype == torch.bfloat16 and C % 8 == 0
    ws = torch.zeros(B * C * 4, device=x.device, dtype=torch.int64) if fast else None
    if stats is not None:
        assert stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() == B * C * 2
    with _Timed(0.0, 5 * _nb(x) + _nb(out)):
        check(lib.c2d_group_norm_bwd(x.data_ptr(), dy.data_ptr(), _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta")), _ptr(add),
                                     out.data_ptr(), _ptr(ws), _ptr(stats) if fast else None, B, HW, C, groups, float(eps),
                                     int(bool(silu)), _dt(x), _stream()),
              "group_norm_bwd")
    return out


def layer_norm_bwd(x, dy, gamma, eps=1e-5, *, add=None, out=None):
    _dev(x)
    C = x.shape[-1]
    M = x.numel() // C
    assert x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype
    assert add is None or (add.is_contiguous() and add.shape == x.shape and add.dtype == x.dtype)
    if out is None:
        out = torch.empty_like(x)
    with _Timed(0.0, 4 * _nb(x) + _nb(out)):
        check(lib.c2d_layer_norm_bwd(x.data_ptr(), dy.data_ptr(), _ptr(_f32(gamma, "gamma")), _ptr(add), out.data_ptr(), M, C,
                                     float(eps), _dt(x), _stream()), "layer_norm_bwd")
    return out


def geglu_bwd(ag, dy, *, out=None):
    _dev(ag)
    F = dy.shape[-1]
    M = dy.numel() // F
    assert ag.is_contiguous() and dy.is_contiguous() and ag.shape[-1] == 2 * F and ag.dtype == dy.dtype
    if out is None:
        out = torch.empty_like(ag)
    with _Timed(0.0, _nb(ag, dy, out)):
        check(lib.c2d_geglu_bwd(ag.data_ptr(), dy.data_ptr(), out.data_ptr(), M, F, _dt(ag), _stream()), "geglu_bwd")
    return out


def attention_bwd(q, k, v, o, dout, heads: int, dq, dk, dv, *, scale: Optional[float] = None, lse=None):
    """Adjoint of ops.attention.  q [B,Nq,h*d], k / v [B,Nkv,h*d], o / dout [B,Nq,h*d]; dq / dk / dv are caller-provided
    (possibly strided) views with the same logical shapes as q / k / v.  lse: the forward pass's log-sum-exp
    (ops.attention(..., lse=buf) that reported `written`): the tensor-core kernels skip their first sweep over the keys."""
    _dev(q)
    B, Nq, C = q.shape
    Nkv = k.shape[1]
    d = C // heads
    for t in (q, k, v, o, dout, dq, dk, dv):
        assert t.stride(2) == 1 and t.dtype == q.dtype
    if scale is None:
        scale = d ** -0.5
    ws = torch.empty(2, B, heads, Nq, device=q.device, dtype=torch.float32)
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and tuple(lse.shape) == (B, heads, Nq)
    lse_ptr = ws[0].data_ptr() if lse is None else lse.data_ptr()
    with _Timed(10.0 * B * heads * Nq * Nkv * d, _nb(q, k, v, o, dout, dq, dk, dv)):
        check(lib.c2d_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), dout.data_ptr(), dq.data_ptr(),
                                    dk.data_ptr(), dv.data_ptr(), lse_ptr, ws[1].data_ptr(), B, heads, Nq, Nkv, d,
                                    q.stride(1), k.stride(1), v.stride(1), o.stride(1), dout.stride(1), dq.stride(1), dk.stride(1),
                                    dv.stride(1), q.stride(0), k.stride(0), v.stride(0), o.stride(0), dout.stride(0), dq.stride(0),
                                    dk.stride(0), dv.stride(0), float(scale), int(lse is not None), _dt(q), _stream()),
              "attention_bwd")
    return dq, dk, dv


def zero_insert2x(x):
    _dev(x)
    B, H, W, C = x.shape
    assert x.is_contiguous()
    z = torch.empty(B, 2 * H, 2 * W, C, device=x.device, dtype=x.dtype)
    with _Timed(0.0, _nb(x, z)):
        check(lib.c2d_zero_insert2x(x.data_ptr(), z.data_ptr(), B, H, W, C, _dt(x), _stream()), "zero_insert2x")
    return z


def sumpool2x2(x):
    _dev(x)
    B, H2, W2, C = x.shape
    assert x.is_contiguous() and H2 % 2 == 0 and W2 % 2 == 0
    y = torch.empty(B, H2 // 2, W2 // 2, C, device=x.device, dtype=x.dtype)
    with _Timed(0.0, _nb(x, y)):
        check(lib.c2d_sumpool2x2(x.data_ptr(), y.data_ptr(), B, H2 // 2, W2 // 2, C, _dt(x), _stream()), "sumpool2x2")
    return y


def slice_channels(x, c0: int, cs: int, *, add=None):
    """y[..., :cs] = x[..., c0:c0+cs] (+ add), contiguous."""
    _dev(x)
    C = x.shape[-1]
    rows = x.numel() // C
    assert x.is_contiguous() and (add is None or (add.is_contiguous() and add.numel() == rows * cs and add.dtype == x.dtype))
    y = torch.empty(*x.shape[:-1], cs, device=x.device, dtype=x.dtype)
    with _Timed(0.0, _nb(y) * 2):
        check(lib.c2d_slice_channels(x.data_ptr(), _ptr(add), y.data_ptr(), rows, C, int(c0), int(cs), _dt(x), _stream()), "slice_channels")
    return y


def mse_loss_grad(pred_nhwc, target_nchw, weight: float, loss_acc):
    """pred [B,H,W,C] (engine dtype), target fp32 [B,C,H,W]; loss_acc: float64 [1] device accumulator (caller-zeroed).
    Returns grad [B,H,W,C] (engine dtype) of weight * mse."""
    _dev(pred_nhwc)
    B, H, W, C = pred_nhwc.shape
    assert pred_nhwc.is_contiguous() and target_nchw.is_contiguous() and target_nchw.dtype == torch.float32
    assert tuple(target_nchw.shape) == (B, C, H, W) and loss_acc.dtype == torch.float64
    grad = torch.empty_like(pred_nhwc)
    check(lib.c2d_mse_loss_grad(pred_nhwc.data_ptr(), target_nchw.data_ptr(), grad.data_ptr(), loss_acc.data_ptr(), B, H * W, C,
                                float(weight), _dt(pred_nhwc), _stream()), "mse_loss_grad")
    return grad


def colsum(x, *, out=None, accumulate: bool = False):
    """x [B,R,C] -> fp32 [B,C] sums over R."""
    _dev(x)
    B, R, C = x.shape
    assert x.is_contiguous()
    if out is None:
        out = torch.empty(B, C, device=x.device, dtype=torch.float32)
        accumulate = False
    check(lib.c2d_colsum(x.data_ptr(), out.data_ptr(), B, R, C, int(accumulate), _dt(x), _stream()), "colsum")
    return out


def gate_bwd(s, af, alpha, dalpha):
    _dev(s)
    assert s.dtype == torch.float32 and af.dtype == torch.float32 and s.is_contiguous() and af.is_contiguous() and s.shape == af.shape
    daf = torch.empty_like(s)
    check(lib.c2d_gate_bwd(s.data_ptr(), af.data_ptr(), _ptr(_f32(alpha, "alpha")), daf.data_ptr(), dalpha.data_ptr(), s.numel(), _stream()), "gate_bwd")
    return daf


def gelu_bwd_bcast(z, dh, K: int):
    _dev(z)
    rows, H = z.shape
    assert z.dtype == torch.float32 and dh.dtype == torch.float32 and z.is_contiguous() and dh.is_contiguous()
    assert tuple(dh.shape) == (rows // K, H)
    dz = torch.empty_like(z)
    check(lib.c2d_gelu_bwd_bcast(z.data_ptr(), dh.data_ptr(), dz.data_ptr(), rows, H, int(K), _stream()), "gelu_bwd_bcast")
    return dz


def sumsq(x, acc):
    _dev(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and acc.dtype == torch.float64
    check(lib.c2d_sumsq(x.data_ptr(), x.numel(), acc.data_ptr(), _stream()), "sumsq")
    return acc


def clip_scale(sumsq_acc, max_norm: float, scale_out, norm_out=None):
    check(lib.c2d_clip_scale(sumsq_acc.data_ptr(), float(max_norm), scale_out.data_ptr(), _ptr(norm_out), _stream()), "clip_scale")
    return scale_out


def adamw_step_sched(param, grad, exp_avg, exp_avg_sq, sched, step_dev, beta1, beta2, eps, weight_decay, grad_scale=None):
    """AdamW with the schedule on the device: sched fp32 [T, 3] = (lr, 1 - beta1^(t+1), 1 - beta2^(t+1)), step_dev int32 [1] = steps
    taken so far (the caller increments it).  No per-step kernel argument: the launch can be replayed from a CUDA graph."""
    _dev(param)
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    assert sched.dtype == torch.float32 and sched.is_contiguous() and sched.dim() == 2 and sched.shape[1] == 3
    assert step_dev.dtype == torch.int32 and step_dev.numel() == 1
    check(lib.c2d_adamw_step_sched(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
                                   sched.data_ptr(), sched.shape[0], step_dev.data_ptr(), float(beta1), float(beta2), float(eps),
                                   float(weight_decay), _ptr(grad_scale), _stream()), "adamw_step_sched")
    return param


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step: int, grad_scale=None):
    _dev(param)
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    check(lib.c2d_adamw_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(), float(lr),
                             float(beta1), float(beta2), float(eps), float(weight_decay), int(step), _ptr(grad_scale), _stream()), "adamw_step")
    return param
