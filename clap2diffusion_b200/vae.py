"""AutoencoderKL decoder of SD-1.5 (SURVEY.md App. A, row 15 of §8a) on libc2d kernels.

Needed for the decoded-image PSNR criterion and for true images/s.  Same engine conventions as
``unet.py``: channels-last activations, packed 3x3 weights, GroupNorm+SiLU fused, shortcut folded into
conv2's residual.  The single-head mid-block attention (4096 tokens, d = 512) goes through the
tensor-core GEMM twice (QK^T, PV) with a row-softmax kernel in between.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops

VAE_BLOCK_OUT = (128, 256, 512, 512)
VAE_SCALING = 0.18215
GROUPS = 32
EPS = 1e-6


class VAEDecoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", dtype=torch.bfloat16,
                 impl: int = ops.IMPL_AUTO):
        self.device, self.dtype, self.impl = torch.device(device), dtype, impl
        self._sd = state_dict
        self.w: Dict[str, torch.Tensor] = {}
        self.fused_gn = dtype == torch.bfloat16       # GroupNorm statistics from the producing tcgen05 epilogues
        self._pack()
        self._sd = None

    # ------------------------------------------------------------------ weights
    def _dev32(self, name):
        return self._sd[name].detach().to(self.device, torch.float32).contiguous()

    def _cast(self, w):
        return w if self.dtype == torch.float32 else ops.cast(w.contiguous(), self.dtype)

    def _conv(self, p):
        self.w[f"{p}.weight"] = ops.pack_conv3x3(self._dev32(f"{p}.weight"), self.dtype)
        self.w[f"{p}.bias"] = self._dev32(f"{p}.bias")

    def _lin(self, p):
        w = self._dev32(f"{p}.weight")
        self.w[f"{p}.weight"] = self._cast(w.reshape(w.shape[0], -1))
        self.w[f"{p}.bias"] = self._dev32(f"{p}.bias")

    def _norm(self, p):
        self.w[f"{p}.weight"], self.w[f"{p}.bias"] = self._dev32(f"{p}.weight"), self._dev32(f"{p}.bias")

    def _resnet(self, p, cin, cout):
        self._norm(f"{p}.norm1"); self._conv(f"{p}.conv1"); self._norm(f"{p}.norm2"); self._conv(f"{p}.conv2")
        if cin != cout:
            self._lin(f"{p}.conv_shortcut")

    def _pack(self):
        # post_quant_conv (1x1, 4->4) with the 1/0.18215 latent scaling folded into its weight
        w = self._dev32("post_quant_conv.weight").reshape(4, 4) / VAE_SCALING
        self.w["post_quant_conv.weight"] = self._cast(w)
        self.w["post_quant_conv.bias"] = self._dev32("post_quant_conv.bias")
        c = VAE_BLOCK_OUT[-1]
        self._conv("decoder.conv_in")
        self._resnet("decoder.mid_block.resnets.0", c, c)
        self._pack_attention("decoder.mid_block.attentions.0")
        self._resnet("decoder.mid_block.resnets.1", c, c)
        prev = c
        for i, cout in enumerate(reversed(VAE_BLOCK_OUT)):
            for j in range(3):
                self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout)
            if i < 3:
                self._conv(f"decoder.up_blocks.{i}.upsamplers.0.conv")
            prev = cout
        self._norm("decoder.conv_norm_out")
        self._conv("decoder.conv_out")

    # ------------------------------------------------------------------ forward
    # Above this many pixels per image the epilogue statistics are not used: thousands of M tiles per image would
    # all hit the same B x C x 2 accumulators (measured: the 256^2 / 512^2 convolutions lose more to atomic
    # contention than the saved statistics pass is worth), so those levels keep the two-pass GroupNorm.
    FUSED_GN_MAX_HW = 128 * 128

    def _stats(self, B: int, C: int, hw: int = 0):
        if not self.fused_gn or hw > self.FUSED_GN_MAX_HW:
            return None
        return torch.zeros(B * C * 2, device=self.device, dtype=torch.int64)

    def _gn(self, x, xs, p, silu=True):
        w = self.w
        if xs is None and self.fused_gn:      # high-resolution level: stand-alone statistics pass, same one-pass apply
            B, C = x.shape[0], x.shape[-1]
            xs = ops.channel_stats(x.view(B, -1, C), torch.zeros(B * C * 2, device=self.device, dtype=torch.int64))
        if xs is not None:
            return ops.group_norm_apply(x, xs, w[f"{p}.weight"], w[f"{p}.bias"], GROUPS, EPS, silu)
        return ops.group_norm(x, w[f"{p}.weight"], w[f"{p}.bias"], GROUPS, EPS, silu)

    def _res(self, p, x, xs):
        """ResnetBlock2D; (x, xs) = activation and its channel statistics (None on the fp32 path); returns the same pair."""
        w = self.w
        B, cout = x.shape[0], w[f"{p}.conv1.weight"].shape[0]
        h = self._gn(x, xs, f"{p}.norm1")
        hw = x.shape[1] * x.shape[2]
        s1 = self._stats(B, cout, hw)
        h = ops.conv3x3(h, w[f"{p}.conv1.weight"], w[f"{p}.conv1.bias"], impl=self.impl, stats=s1)
        h = self._gn(h, s1, f"{p}.norm2")
        sc = x
        if f"{p}.conv_shortcut.weight" in w:
            sc = ops.linear(x, w[f"{p}.conv_shortcut.weight"], w[f"{p}.conv_shortcut.bias"], impl=self.impl)
        so = self._stats(B, cout, hw)
        return ops.conv3x3(h, w[f"{p}.conv2.weight"], w[f"{p}.conv2.bias"], residual=sc, impl=self.impl, stats=so), so

    # The published SD-1.5 VAE files carry the attention block under its pre-0.19 diffusers names (query / key / value /
    # proj_attn, 1x1-conv or linear shaped); diffusers renames them on load (_convert_deprecated_attention_blocks).
    _ATTN_ALIASES = {"to_q": ("to_q", "query"), "to_k": ("to_k", "key"), "to_v": ("to_v", "value"),
                     "to_out.0": ("to_out.0", "proj_attn")}

    def _attn_param(self, a, name, kind):
        for alias in self._ATTN_ALIASES[name]:
            if f"{a}.{alias}.{kind}" in self._sd:
                t = self._dev32(f"{a}.{alias}.{kind}")
                return t.reshape(t.shape[0], -1) if kind == "weight" else t
        have = sorted(k for k in self._sd if k.startswith(a + "."))
        raise KeyError(f"VAE attention block {a!r}: no {kind} for {' / '.join(self._ATTN_ALIASES[name])} in the state dict "
                       f"(keys under the block: {have})")

    def _pack_attention(self, a):
        self._norm(f"{a}.group_norm")
        wq, wk, wv = (self._attn_param(a, n, "weight") for n in ("to_q", "to_k", "to_v"))
        self.w[f"{a}.qkv.weight"] = self._cast(torch.cat([wq, wk, wv], 0))
        self.w[f"{a}.qkv.bias"] = torch.cat([self._attn_param(a, n, "bias") for n in ("to_q", "to_k", "to_v")], 0).contiguous()
        self.w[f"{a}.to_out.0.weight"] = self._cast(self._attn_param(a, "to_out.0", "weight"))
        self.w[f"{a}.to_out.0.bias"] = self._attn_param(a, "to_out.0", "bias")

    def _mid_attention(self, x, xs, a="decoder.mid_block.attentions.0"):
        w = self.w
        B, H, W, C = x.shape
        N = H * W
        xf = x.view(B, N, C)
        hn = self._gn(xf, xs, f"{a}.group_norm", silu=False)
        qkv = ops.linear(hn, w[f"{a}.qkv.weight"], w[f"{a}.qkv.bias"], impl=self.impl)          # [B,N,3C]
        o = torch.empty(B, N, C, device=x.device, dtype=x.dtype)
        for b in range(B):           # single head, d = C = 512: materialised scores per image
            q, k, v = qkv[b, :, :C], qkv[b, :, C:2 * C], qkv[b, :, 2 * C:]
            s = ops.linear(q, k.contiguous(), impl=self.impl)                                    # q k^T  [N,N]
            p = ops.softmax_rows(s, scale=C ** -0.5)
            vt = ops.transpose(v.contiguous().view(1, N, C)).view(C, N)                          # [C,N]
            ops.linear(p, vt, out=o[b], impl=self.impl)                                          # p v    [N,C]
        so = self._stats(B, C)
        out = ops.linear(o, w[f"{a}.to_out.0.weight"], w[f"{a}.to_out.0.bias"], residual=xf, impl=self.impl,
                         stats=so, stats_rows=N)
        return out.view(B, H, W, C), so

    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """z fp32 NCHW [B,4,h,w] (scaled latents) -> image fp32 NCHW [B,3,8h,8w]."""
        w = self.w
        x = ops.nchw_to_nhwc(z.contiguous().float(), self.dtype)
        B = x.shape[0]
        h = ops.linear(x, w["post_quant_conv.weight"], w["post_quant_conv.bias"], impl=self.impl)
        h = ops.conv3x3(h, w["decoder.conv_in.weight"], w["decoder.conv_in.bias"], impl=self.impl)
        hs = ops.channel_stats(h, self._stats(B, h.shape[-1])) if self.fused_gn else None
        h, hs = self._res("decoder.mid_block.resnets.0", h, hs)
        h, hs = self._mid_attention(h, hs)
        h, hs = self._res("decoder.mid_block.resnets.1", h, hs)
        for i in range(4):
            for j in range(3):
                h, hs = self._res(f"decoder.up_blocks.{i}.resnets.{j}", h, hs)
            if i < 3:
                n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                if self.dtype == torch.bfloat16 or self.fused_gn:
                    hs = self._stats(B, w[f"{n}.weight"].shape[0], 4 * h.shape[1] * h.shape[2])
                    h = ops.conv3x3(ops.upsample2x(h), w[f"{n}.weight"], w[f"{n}.bias"], impl=self.impl, stats=hs)
                else:
                    h = ops.conv3x3(h, w[f"{n}.weight"], w[f"{n}.bias"], upsample=True, impl=self.impl)
        h = self._gn(h, hs, "decoder.conv_norm_out")
        h = ops.conv3x3(h, w["decoder.conv_out.weight"], w["decoder.conv_out.bias"], impl=self.impl)
        return ops.nhwc_to_nchw(h)


class VAEEncoder(VAEDecoder):
    """AutoencoderKL encoder + quant_conv of SD-1.5 (latents (4, 64, 64) as stored by the reference's
    data/audiocaps_latent_v4.py:185) on the same kernels as the decoder.  The three Downsample2D(padding=0) layers are
    `c2d_conv3x3_down` (zero padding on the right / bottom only: TMA out-of-bounds fill at an un-shifted window)."""

    def _pack(self):
        self._conv_any("encoder.conv_in")
        prev = VAE_BLOCK_OUT[0]
        for i, cout in enumerate(VAE_BLOCK_OUT):
            for j in range(2):
                self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout)
            if i < 3:
                self._conv(f"encoder.down_blocks.{i}.downsamplers.0.conv")
            prev = cout
        c = VAE_BLOCK_OUT[-1]
        self._resnet("encoder.mid_block.resnets.0", c, c)
        self._pack_attention("encoder.mid_block.attentions.0")
        self._resnet("encoder.mid_block.resnets.1", c, c)
        self._norm("encoder.conv_norm_out")
        self._conv("encoder.conv_out")
        self._lin("quant_conv")
        # encode(): posterior mode with the 0.18215 latent scaling folded into the four mean rows of quant_conv
        wq = self._dev32("quant_conv.weight").reshape(8, 8)[:4] * VAE_SCALING
        self.w["quant_conv.mean_scaled.weight"] = self._cast(wq.contiguous())
        self.w["quant_conv.mean_scaled.bias"] = (self._dev32("quant_conv.bias")[:4] * VAE_SCALING).contiguous()

    def _conv_any(self, p):
        self._conv(p)

    @torch.no_grad()
    def _trunk(self, img: torch.Tensor) -> torch.Tensor:
        """img fp32 NCHW [B,3,H,W] -> encoder.conv_out activations, NHWC [B,H/8,W/8,8]."""
        w = self.w
        x = ops.nchw_to_nhwc(img.contiguous().float(), self.dtype)
        B = x.shape[0]
        h = ops.conv3x3(x, w["encoder.conv_in.weight"], w["encoder.conv_in.bias"], impl=self.impl)
        hs = None            # 512^2 / 256^2 levels: stand-alone statistics pass inside _gn (see FUSED_GN_MAX_HW)
        for i in range(4):
            for j in range(2):
                h, hs = self._res(f"encoder.down_blocks.{i}.resnets.{j}", h, hs)
            if i < 3:
                n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                hs = self._stats(B, w[f"{n}.weight"].shape[0], (h.shape[1] // 2) * (h.shape[2] // 2))
                h = ops.conv3x3_down(h, w[f"{n}.weight"], w[f"{n}.bias"], impl=self.impl, stats=hs)
        h, hs = self._res("encoder.mid_block.resnets.0", h, hs)
        h, hs = self._mid_attention(h, hs, "encoder.mid_block.attentions.0")
        h, hs = self._res("encoder.mid_block.resnets.1", h, hs)
        h = self._gn(h, hs, "encoder.conv_norm_out")
        return ops.conv3x3(h, w["encoder.conv_out.weight"], w["encoder.conv_out.bias"], impl=self.impl)

    @torch.no_grad()
    def moments(self, img: torch.Tensor):
        """(mean, logvar) of the posterior, fp32 NCHW [B,4,H/8,W/8] each (logvar clamped to [-30, 20] like diffusers'
        DiagonalGaussianDistribution; the clamp of this small diagnostic tensor is the one torch elementwise call here)."""
        h = self._trunk(img)
        m = ops.nhwc_to_nchw(ops.linear(h, self.w["quant_conv.weight"], self.w["quant_conv.bias"], impl=self.impl))
        return m[:, :4].contiguous(), m[:, 4:].clamp(-30.0, 20.0).contiguous()

    @torch.no_grad()
    def encode(self, img: torch.Tensor) -> torch.Tensor:
        """Scaled latents (posterior mode x 0.18215), fp32 NCHW [B,4,H/8,W/8]: what VAEDecoder.decode takes."""
        h = self._trunk(img)
        z = ops.linear(h, self.w["quant_conv.mean_scaled.weight"], self.w["quant_conv.mean_scaled.bias"], impl=self.impl)
        return ops.nhwc_to_nchw(z)


def encoder_param_shapes() -> Dict[str, tuple]:
    """diffusers state-dict names -> shapes of the AutoencoderKL encoder + quant_conv (34,163,664)."""
    out: Dict[str, tuple] = {}

    def conv(n, cin, cout, k):
        out[f"{n}.weight"] = (cout, cin, k, k)
        out[f"{n}.bias"] = (cout,)

    def norm(n, c):
        out[f"{n}.weight"] = (c,)
        out[f"{n}.bias"] = (c,)

    def resnet(n, cin, cout):
        norm(f"{n}.norm1", cin); conv(f"{n}.conv1", cin, cout, 3); norm(f"{n}.norm2", cout); conv(f"{n}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{n}.conv_shortcut", cin, cout, 1)

    conv("encoder.conv_in", 3, VAE_BLOCK_OUT[0], 3)
    prev = VAE_BLOCK_OUT[0]
    for i, cout in enumerate(VAE_BLOCK_OUT):
        for j in range(2):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout)
        if i < 3:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        prev = cout
    c = VAE_BLOCK_OUT[-1]
    resnet("encoder.mid_block.resnets.0", c, c)
    a = "encoder.mid_block.attentions.0"
    norm(f"{a}.group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out[f"{a}.{n}.weight"] = (c, c)
        out[f"{a}.{n}.bias"] = (c,)
    resnet("encoder.mid_block.resnets.1", c, c)
    norm("encoder.conv_norm_out", c); conv("encoder.conv_out", c, 8, 3); conv("quant_conv", 8, 8, 1)
    return out


def param_shapes() -> Dict[str, tuple]:
    """diffusers state-dict names -> shapes of the AutoencoderKL decoder + post_quant_conv (49,490,199)."""
    out: Dict[str, tuple] = {}

    def conv(n, cin, cout, k):
        out[f"{n}.weight"] = (cout, cin, k, k)
        out[f"{n}.bias"] = (cout,)

    def norm(n, c):
        out[f"{n}.weight"] = (c,)
        out[f"{n}.bias"] = (c,)

    def resnet(n, cin, cout):
        norm(f"{n}.norm1", cin); conv(f"{n}.conv1", cin, cout, 3); norm(f"{n}.norm2", cout); conv(f"{n}.conv2", cout, cout, 3)
        if cin != cout:
            conv(f"{n}.conv_shortcut", cin, cout, 1)

    c = VAE_BLOCK_OUT[-1]
    conv("post_quant_conv", 4, 4, 1); conv("decoder.conv_in", 4, c, 3)
    resnet("decoder.mid_block.resnets.0", c, c)
    a = "decoder.mid_block.attentions.0"
    norm(f"{a}.group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        out[f"{a}.{n}.weight"] = (c, c)
        out[f"{a}.{n}.bias"] = (c,)
    resnet("decoder.mid_block.resnets.1", c, c)
    prev = c
    for i, cout in enumerate(reversed(VAE_BLOCK_OUT)):
        for j in range(3):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout)
        if i < 3:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        prev = cout
    norm("decoder.conv_norm_out", VAE_BLOCK_OUT[0]); conv("decoder.conv_out", VAE_BLOCK_OUT[0], 3, 3)
    return out
