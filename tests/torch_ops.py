"""TEST DOUBLE for ``clap2diffusion_b200.ops``: the same function signatures implemented with plain torch,
so the HOST logic of the product (weight packing, topology, skip wiring, hoisting, processor protocol,
drop-in module plumbing) can be checked against the oracle on a machine without a GPU.

It is installed only by tests (``with torch_ops.installed(): ...``) and lives under tests/; the product
never imports it and has no CPU path of its own.
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as F

from clap2diffusion_b200 import ops as real_ops

ACT = {0: lambda v: v, 1: F.gelu, 2: F.silu, 3: F.relu}


STATS_SCALE = float(1 << 20)


def _add_stats(stats, y, B):
    """Mirror of the epilogue statistics: per-channel (sum, sumsq) as 2^20 fixed-point int64 [B, C, 2]."""
    C = y.shape[-1]
    yy = y.float().reshape(B, -1, C)
    st = torch.stack([yy.sum(1), (yy * yy).sum(1)], -1)
    stats.view(B, C, 2).add_(torch.round(st.double() * STATS_SCALE).to(torch.int64))


def _add_row_stats(row_stats, y):
    yy = y.float().reshape(-1, y.shape[-1])
    st = torch.stack([yy.sum(1), (yy * yy).sum(1)], -1)
    row_stats.view(-1, 2).add_(torch.round(st.double() * STATS_SCALE).to(torch.int64))


def _ln_coeffs(ln_stats, K, eps):
    st = ln_stats.view(-1, 2).double() / STATS_SCALE / K
    mean = st[:, 0]
    var = (st[:, 1] - mean * mean).clamp_min(0.0)
    rstd = 1.0 / torch.sqrt(var + eps)
    return mean.float(), rstd.float()


def pack_lnfold(w, gamma, beta, bias, dtype, eps=1e-5, out_dtype=None):
    wg = w * gamma[None, :]
    cs = wg.to(torch.bfloat16).float().sum(1)
    bo = (w * beta[None, :]).sum(1) + (bias if bias is not None else 0.0)
    return real_ops.LNFold(wg.to(out_dtype or dtype).contiguous(), cs.contiguous(), bo.contiguous(), eps)


def linear(x, w, bias=None, *, act=0, residual=None, rowvec=None, rows_per_vec=1, out=None, impl=0, x2=None,
           stats=None, stats_rows=0, row_stats=None, ln=None, ln_stats=None):
    if x2 is not None:
        x = torch.cat([x, x2], -1)
    if ln is not None:
        mean, rstd = _ln_coeffs(ln_stats, x.shape[-1], ln.eps)
        acc = F.linear(x.float(), ln.w.float()).reshape(-1, ln.w.shape[0])
        y = (rstd[:, None] * (acc - mean[:, None] * ln.colsum[None, :]) + ln.bias[None, :]).reshape(*x.shape[:-1], -1)
    else:
        y = F.linear(x.float(), w.float(), None if bias is None else bias.float())
    if rowvec is not None:
        M = y.numel() // y.shape[-1]
        idx = torch.arange(M, device=y.device) // rows_per_vec
        y = (y.reshape(M, -1) + rowvec.reshape(-1, y.shape[-1])[idx]).reshape(y.shape)
    y = ACT[act](y)
    if residual is not None:
        y = y + residual.float().reshape(y.shape)
    if stats is not None:
        _add_stats(stats, y, (y.numel() // y.shape[-1]) // stats_rows)
    if row_stats is not None:
        _add_row_stats(row_stats, y)
    y = y.to(x.dtype)
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y


def pack_geglu(w, bias, dtype):
    F2, K = w.shape
    Fh = F2 // 2
    idx = []
    for blk in range(Fh // 64):
        idx += list(range(blk * 64, blk * 64 + 64)) + list(range(Fh + blk * 64, Fh + blk * 64 + 64))
    idx = torch.tensor(idx, device=w.device)
    return w[idx].to(dtype).contiguous(), (None if bias is None else bias[idx].float().contiguous())


def geglu_linear(x, w_packed, bias_packed, *, out=None, impl=0, ln=None, ln_stats=None):
    if ln is not None:
        mean, rstd = _ln_coeffs(ln_stats, x.shape[-1], ln.eps)
        acc = F.linear(x.float(), ln.w.float()).reshape(-1, ln.w.shape[0])
        y = (rstd[:, None] * (acc - mean[:, None] * ln.colsum[None, :]) + ln.bias[None, :]).reshape(*x.shape[:-1], -1)
    else:
        y = F.linear(x.float(), w_packed.float(), bias_packed)
    M = y.numel() // y.shape[-1]
    y = y.reshape(M, -1, 2, 64)                     # blocks of (a[64], g[64])
    r = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(*x.shape[:-1], -1)
    return r.to(x.dtype)


def geglu(x, *, out=None):
    a, g = x.float().chunk(2, dim=-1)
    return (a * F.gelu(g)).to(x.dtype)


def pack_conv3x3(w, dtype):
    return w.permute(0, 2, 3, 1).contiguous().to(dtype)


def conv3x3(x, w_packed, bias=None, *, rowvec=None, residual=None, stride=1, upsample=False, out=None, impl=0,
            stats=None):
    xin = x.float().permute(0, 3, 1, 2)
    if upsample:
        xin = F.interpolate(xin, scale_factor=2.0, mode="nearest")
    w = w_packed.float().permute(0, 3, 1, 2)
    y = F.conv2d(xin, w, None if bias is None else bias.float(), stride=stride, padding=1)
    if rowvec is not None:
        y = y + rowvec.reshape(y.shape[0], -1)[:, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual.float().reshape(y.shape)
    if stats is not None:
        _add_stats(stats, y, y.shape[0])
    return y.contiguous().to(x.dtype)


def conv3x3_down(x, w_packed, bias=None, *, out=None, impl=0, stats=None):
    w = w_packed.float().permute(0, 3, 1, 2)
    xin = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    y = F.conv2d(xin, w, None if bias is None else bias.float(), stride=2, padding=0).permute(0, 2, 3, 1)
    if stats is not None:
        _add_stats(stats, y, y.shape[0])
    return y.to(x.dtype).contiguous()


def group_norm(x, gamma, beta, groups=32, eps=1e-5, silu=False, *, x2=None, raw_cat=None, out=None):
    xx = x if x2 is None else torch.cat([x, x2], dim=-1)
    if raw_cat is not None:
        raw_cat.copy_(xx)
    B, C = xx.shape[0], xx.shape[-1]
    y = F.group_norm(xx.float().reshape(B, -1, C).transpose(1, 2), groups, gamma, beta, eps).transpose(1, 2)
    if silu:
        y = F.silu(y)
    return y.reshape(xx.shape).contiguous().to(x.dtype)


def channel_stats(x, stats):
    _add_stats(stats, x, x.shape[0])
    return stats


def group_norm_apply(x, stats, gamma, beta, groups=32, eps=1e-5, silu=False, *, x2=None, stats2=None, out=None):
    xx = x if x2 is None else torch.cat([x, x2], dim=-1)
    B, C = xx.shape[0], xx.shape[-1]
    st = stats.view(B, -1, 2) if stats2 is None else torch.cat([stats.view(B, -1, 2), stats2.view(B, -1, 2)], 1)
    n = (xx.numel() // (B * C)) * (C // groups)
    g = st.view(B, groups, C // groups, 2).sum(2).double() / STATS_SCALE / n
    mean, var = g[..., 0], (g[..., 1] - g[..., 0] ** 2).clamp_min(0.0)
    rstd = 1.0 / torch.sqrt(var + eps)
    cm = mean.float().repeat_interleave(C // groups, 1)[:, None, :]
    cr = rstd.float().repeat_interleave(C // groups, 1)[:, None, :]
    y = (xx.float().reshape(B, -1, C) - cm) * cr * gamma + beta
    if silu:
        y = F.silu(y)
    return y.reshape(xx.shape).contiguous().to(x.dtype)


def layer_norm(x, gamma, beta, eps=1e-5, *, out=None):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps).to(x.dtype)


def attention(q, k, v, heads, *, scale=None, mask=None, out=None, impl=0, lse=None):
    if lse is not None:          # the double never produces the by-product: the adjoint recomputes it
        return attention(q, k, v, heads, scale=scale, mask=mask, out=out, impl=impl), False
    B, Nq, C = q.shape
    d = C // heads
    scale = d ** -0.5 if scale is None else scale
    qh = q.float().reshape(B, Nq, heads, d).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.float().reshape(B, -1, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * scale
    if mask is not None:
        s = s.masked_fill(~mask.bool()[:, None, None, :], -torch.finfo(s.dtype).max)
    o = (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Nq, C)
    return o.to(q.dtype)


def xattn_supported(x, heads, T, T2=0):
    return True


def xattn_packable(C, heads, T, dtype, T2=0):
    return True


class _KV(real_ops.XattnKV):
    def __init__(self, kv, kv2, heads, lambda2=1.0):
        super().__init__(None, kv.shape[0], kv.shape[2] // 2, heads, kv.shape[1], 0 if kv2 is None else kv2.shape[1], kv, lambda2)
        self.kv, self.kv2, self.heads = kv, kv2, heads
        self.B, self.T, self.C = kv.shape[0], kv.shape[1], kv.shape[2] // 2
        self.T2 = 0 if kv2 is None else kv2.shape[1]


def xattn_pack_kv(kv, heads, kv2=None, lambda2=1.0):
    return _KV(kv, kv2, heads, lambda2)


def xattn(x, kvp, *, wq=None, q_bias=None, ln=None, ln_stats=None, scale=None, lambda2=None, out=None):
    """q is rounded to the activation dtype (it is the bf16 A operand of the score MMA in the kernel)."""
    C = x.shape[-1]
    if ln is not None:
        q = linear(x, None, ln=ln, ln_stats=ln_stats)
    else:
        q = linear(x, wq, q_bias)
    lambda2 = kvp.lambda2 if lambda2 is None else lambda2
    o = _attn_probs_rounded(q, kvp.kv[..., :C], kvp.kv[..., C:], kvp.heads, scale)
    if kvp.kv2 is not None:
        o = o + lambda2 * _attn_probs_rounded(q, kvp.kv2[..., :C], kvp.kv2[..., C:], kvp.heads, scale)
    return o.to(x.dtype)


def _attn_probs_rounded(q, k, v, heads, scale):
    B, Nq, C = q.shape
    d = C // heads
    scale = d ** -0.5 if scale is None else scale
    qh = q.float().reshape(B, Nq, heads, d).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.float().reshape(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, -1)
    return (p @ vh).transpose(1, 2).reshape(B, Nq, C)


def audio_context(ehs, audio, w1, b1, w2, b2, alpha, mode, *, out=None):
    ap = F.linear(F.gelu(F.linear(audio.float(), w1.float(), b1)), w2.float(), b2)
    if mode == real_ops.AUDIO_ADD:
        return (ehs.float() + torch.sigmoid(alpha) * ap.mean(dim=1, keepdim=True)).to(ehs.dtype)
    if ap.shape[1] > 4:
        ap = F.adaptive_avg_pool1d(ap.transpose(1, 2), 4).transpose(1, 2)
    return torch.cat([ehs.float(), ap], dim=1).to(ehs.dtype)


def timestep_embedding(t, dim=320, *, out=None):
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t.float()[:, None] * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], -1)


def unary(x, act=0, *, out_dtype=None, out=None):
    return ACT[act](x.float()).to(out_dtype or x.dtype)


def cast(x, dtype, *, out=None):
    return x.to(dtype)


def add(a, b, *, out=None):
    return a + b


def upsample2x(x, *, out=None):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()


def concat(x1, x2, *, out=None):
    return torch.cat([x1, x2], dim=-1)


def nchw_to_nhwc(x, dtype, *, out=None):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nhwc_to_nchw(x, *, out=None):
    return x.float().permute(0, 3, 1, 2).contiguous()


def cfg_sched_step(eps2, x, xin2, guidance, coef, trace=None):
    B = x.shape[0]
    e = eps2.float().permute(0, 3, 1, 2) if eps2.dim() == 4 else eps2.float().reshape(2 * B, x.shape[2], x.shape[3], 4).permute(0, 3, 1, 2)
    eu, ec = e[:B], e[B:]
    x.copy_(coef[0] * x + coef[1] * (eu + guidance * (ec - eu)))
    if trace is not None:
        trace.copy_(x)
    xi = (x * coef[2]).permute(0, 2, 3, 1).to(xin2.dtype)
    xin2[..., :4].copy_(torch.cat([xi, xi], 0).reshape(*xin2.shape[:-1], 4))
    return x


def softmax_rows(x, scale=1.0, *, out=None):
    return torch.softmax(x.float() * scale, -1).to(x.dtype)


def transpose(x, *, out=None):
    return x.transpose(1, 2).contiguous()


def bcast_add(a, b, B, K, D, a_mode, b_mode, *, out=None):
    def view(t, m):
        return t.reshape(B, K, D) if m == 0 else (t.reshape(B, 1, D) if m == 1 else t.reshape(1, K, D))
    return (view(a, a_mode) + view(b, b_mode)).expand(B, K, D).contiguous()


def hier_assign(tokens, anchors, w1, b1, w2, b2, temperature):
    t = tokens.float()
    sim = torch.einsum("bkd,ld->bkl", F.normalize(t, dim=-1), F.normalize(anchors.float(), dim=-1)) * 10.0
    gate = F.linear(F.gelu(F.linear(t, w1.float(), b1)), w2.float(), b2)
    return torch.softmax((sim + gate) / temperature, dim=-1)


def hier_route(tok10, assign, hw, routing, gates):
    a = assign
    if hw is not None:
        a = a * hw[:, None, :]
        a = a / (a.sum(-1, keepdim=True) + 1e-8)
    r = a @ torch.softmax(routing, dim=1)
    return tuple((tok10.float() * r[:, :, i:i + 1] * torch.sigmoid(gates[i])).to(tok10.dtype) for i in range(3))


def norm_scale(x, target=60.0, per_sample=False, *, out=None):
    n = x.float().norm(dim=-1, keepdim=True)
    m = n.mean(dim=(1, 2), keepdim=True) if per_sample else n.mean()
    return (x.float() * torch.where(m > 0, target / m, torch.ones_like(m))).to(x.dtype)


def legacy_combine(fg, bg, amb, hierarchy_weights, D):
    B = fg.shape[0]
    w = torch.softmax(hierarchy_weights, 0)
    return torch.cat([fg.reshape(B, -1, D) * w[0], bg.reshape(B, -1, D) * w[1], amb.reshape(B, -1, D) * w[2]], 1)


# ---- CLAP audio tower doubles -------------------------------------------------------------------------
def stft_frames(wave, window, hop, n_frames, *, out=None):
    n_fft = window.numel()
    w = F.pad(wave[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    idx = torch.arange(n_frames)[:, None] * hop + torch.arange(n_fft)[None, :]
    return (w[:, idx] * window[None, None, :]).reshape(-1, n_fft).contiguous()


def stft_frames_split(wave, window, hop, n_frames, *, out=None):
    x = stft_frames(wave, window, hop, n_frames)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo, hi], dim=1).contiguous()


def power_spectrum(dft, *, nb=None, im_off=None, out=None):
    nb = dft.shape[1] // 2 if nb is None else nb
    im_off = nb if im_off is None else im_off
    d = dft.float()
    return d[:, :nb] ** 2 + d[:, im_off:im_off + nb] ** 2


def log_mel_affine(x, a, b, floor=1e-10, *, out=None):
    y = 10.0 * torch.log10(torch.clamp(x, min=floor)) * a[None, :] + b[None, :]
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y


def clap_patches(mel, dtype, *, out=None):
    B = mel.shape[0]
    x = F.interpolate(mel[:, None], (1024, 64), mode="bicubic", align_corners=True)
    img = x.reshape(B, 4, 256, 64).permute(0, 1, 3, 2).reshape(B, 256, 256)
    p = img.reshape(B, 64, 4, 64, 4).permute(0, 1, 3, 2, 4).reshape(B * 4096, 16)
    return p.contiguous().to(dtype)


def window_attention(qkv, bias, H, W, heads, shift, *, out=None):
    B, N, C3 = qkv.shape
    C, d = C3 // 3, C3 // 3 // heads
    t = qkv.float().view(B, H, W, C3)
    if shift:
        t = torch.roll(t, (-shift, -shift), (1, 2))
    nh, nw = H // 8, W // 8
    win = t.view(B, nh, 8, nw, 8, C3).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64, C3)
    q, k, v = (win[..., i * C:(i + 1) * C].reshape(-1, 64, heads, d).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) * d ** -0.5 + bias[None]
    if shift:
        img = torch.zeros(1, H, W, 1)
        cnt = 0
        for hs in (slice(0, -8), slice(-8, -shift), slice(-shift, None)):
            for ws in (slice(0, -8), slice(-8, -shift), slice(-shift, None)):
                img[:, hs, ws, :] = cnt
                cnt += 1
        m = img.view(1, nh, 8, nw, 8, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64)
        dm = m[:, None, :] - m[:, :, None]
        dm = torch.where(dm != 0, torch.full_like(dm, -100.0), torch.zeros_like(dm))
        s = (s.view(B, nh * nw, heads, 64, 64) + dm[None, :, None]).view(-1, heads, 64, 64)
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(-1, 64, C)
    o = o.view(B, nh, nw, 8, 8, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o.reshape(B, N, C).to(qkv.dtype)


def patch_merge(x, H, W, *, out=None):
    B, N, C = x.shape
    t = x.view(B, H, W, C)
    return torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1).reshape(B, -1, 4 * C).contiguous()


def token_mean(x, *, out=None):
    return x.float().mean(1)


def l2_normalize(x, eps=1e-12, *, out=None):
    return F.normalize(x, dim=-1, eps=eps)


# ---------------------------------------------------------------------------------------------------------
# backward / optimiser doubles (stage-3 step).  Each one is torch AUTOGRAD of the forward double -- an independent
# route to the numbers the hand-written adjoint kernels must produce.
# ---------------------------------------------------------------------------------------------------------
def _vjp(fn, inputs, gy):
    xs = [t.detach().float().requires_grad_(True) for t in inputs]
    with torch.enable_grad():
        y = fn(*xs)
        return torch.autograd.grad(y, xs, gy.float().reshape(y.shape))


def group_norm_bwd(x, dy, gamma, beta, groups=32, eps=1e-5, silu=False, *, add=None, out=None, stats=None):
    B, C = x.shape[0], x.shape[-1]

    def f(t):
        y = F.group_norm(t.reshape(B, -1, C).transpose(1, 2), groups, gamma.float(), beta.float(), eps).transpose(1, 2)
        return (F.silu(y) if silu else y).reshape(x.shape)
    g = _vjp(f, [x], dy)[0]
    if add is not None:
        g = g + add.float()
    return g.to(x.dtype)


def layer_norm_bwd(x, dy, gamma, eps=1e-5, *, add=None, out=None):
    g = _vjp(lambda t: F.layer_norm(t, (x.shape[-1],), gamma.float(), torch.zeros_like(gamma).float(), eps), [x], dy)[0]
    if add is not None:
        g = g + add.float()
    return g.to(x.dtype)


def geglu_bwd(ag, dy, *, out=None):
    Fh = dy.shape[-1]
    return _vjp(lambda t: t[..., :Fh] * F.gelu(t[..., Fh:]), [ag], dy)[0].to(ag.dtype)


def attention_bwd(q, k, v, o, dout, heads, dq, dk, dv, *, scale=None, lse=None):
    def f(qq, kk, vv):
        B, Nq, C = qq.shape
        d = C // heads
        sc = d ** -0.5 if scale is None else scale
        qh = qq.reshape(B, Nq, heads, d).transpose(1, 2)
        kh = kk.reshape(B, -1, heads, d).transpose(1, 2)
        vh = vv.reshape(B, -1, heads, d).transpose(1, 2)
        return (torch.softmax(qh @ kh.transpose(-1, -2) * sc, -1) @ vh).transpose(1, 2).reshape(B, Nq, C)
    gq, gk, gv = _vjp(f, [q, k, v], dout)
    dq.copy_(gq); dk.copy_(gk); dv.copy_(gv)
    return dq, dk, dv


def zero_insert2x(x):
    B, H, W, C = x.shape
    z = torch.zeros(B, 2 * H, 2 * W, C, dtype=x.dtype)
    z[:, ::2, ::2] = x
    return z


def sumpool2x2(x):
    B, H2, W2, C = x.shape
    return x.float().reshape(B, H2 // 2, 2, W2 // 2, 2, C).sum(dim=(2, 4)).to(x.dtype)


def slice_channels(x, c0, cs, *, add=None):
    y = x[..., c0:c0 + cs].float()
    if add is not None:
        y = y + add.float().reshape(y.shape)
    return y.contiguous().to(x.dtype)


def mse_loss_grad(pred_nhwc, target_nchw, weight, loss_acc):
    p = pred_nhwc.detach().float().requires_grad_(True)
    with torch.enable_grad():
        loss = weight * F.mse_loss(p.permute(0, 3, 1, 2), target_nchw.float())
        g = torch.autograd.grad(loss, p)[0]
    loss_acc += loss.detach().double()
    return g.to(pred_nhwc.dtype)


def colsum(x, *, out=None, accumulate=False):
    s = x.float().sum(1)
    if out is None:
        return s
    if accumulate:
        out += s.reshape(out.shape)
    else:
        out.copy_(s.reshape(out.shape))
    return out


def gate_bwd(s, af, alpha, dalpha):
    gate = torch.sigmoid(alpha.float().reshape(()))
    dalpha += (s * af).sum() * gate * (1 - gate)
    return gate * s


def gelu_bwd_bcast(z, dh, K):
    zz = z.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        g = torch.autograd.grad(F.gelu(zz), zz, dh.repeat_interleave(K, 0) / K)[0]
    return g


def sumsq(x, acc):
    acc += (x.double() ** 2).sum()
    return acc


def clip_scale(sumsq_acc, max_norm, scale_out, norm_out=None):
    norm = float(sumsq_acc.reshape(-1)[0]) ** 0.5
    scale_out.fill_(min(1.0, max_norm / (norm + 1e-6)))
    if norm_out is not None:
        norm_out.fill_(norm)
    return scale_out


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None):
    g = grad * (1.0 if grad_scale is None else float(grad_scale.reshape(-1)[0]))
    param.mul_(1 - lr * weight_decay)
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    param.addcdiv_(exp_avg / (1 - beta1 ** step), (exp_avg_sq / (1 - beta2 ** step)).sqrt() + eps, value=-lr)
    return param


def adamw_step_sched(param, grad, exp_avg, exp_avg_sq, sched, step_dev, beta1, beta2, eps, weight_decay, grad_scale=None):
    t = min(max(int(step_dev.reshape(-1)[0]), 0), sched.shape[0] - 1)
    lr, bc1, bc2 = (float(v) for v in sched[t])
    g = grad * (1.0 if grad_scale is None else float(grad_scale.reshape(-1)[0]))
    param.mul_(1 - lr * weight_decay)
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    param.addcdiv_(exp_avg / bc1, (exp_avg_sq / bc2).sqrt() + eps, value=-lr)
    return param


def require_cuda(x, who):
    """The CPU host-logic tests run the product's launch sequences on torch doubles: no device requirement."""
    return None


_NAMES = [n for n, f in list(globals().items()) if callable(f) and not n.startswith("_") and hasattr(real_ops, n)
          and n not in ("installed",)]


@contextlib.contextmanager
def installed():
    """Swap every libc2d-backed op for its torch double (and lift the CUDA-only guards)."""
    saved = {n: getattr(real_ops, n) for n in _NAMES}
    try:
        for n in _NAMES:
            setattr(real_ops, n, globals()[n])
        yield
    finally:
        for n, f in saved.items():
            setattr(real_ops, n, f)
