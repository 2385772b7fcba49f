"""fp32 PyTorch restatement of the Stable-Diffusion-1.5 UNet, DDIM/Euler schedulers, CFG and
VAE decoder.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates third-party ``diffusers==0.23.1`` (reference ``requirements.txt:7``; base model
``runwayml/stable-diffusion-v1-5`` named at reference ``configs/training_config.yaml:2``), following
SURVEY.md App. A / App. B.  diffusers is not installed and not vendored in the reference, so this
file is PARITY-UNPINNED except for: exact parameter counts (859,520,964 UNet; 49,490,179 + 20 VAE
decoder incl. post_quant_conv), the 16+16 attention-site census and names, and DDIM-50 timesteps
981..1.  State-dict keys follow diffusers' naming so real checkpoints would map 1:1.

Everything is functional: ``fn(sd, ...)`` with ``sd`` a dict name -> fp32 tensor (NCHW convs).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .weights import P, conv, linear, norm

BLOCK_OUT = (320, 640, 1280, 1280)
CROSS_DIM = 768
HEADS = 8
TEMB = 1280
GROUPS = 32


# =======================================================================================
# Parameter spec
# =======================================================================================
def _resnet_spec(name: str, cin: int, cout: int, temb: Optional[int] = TEMB) -> List[P]:
    s = norm(f"{name}.norm1", cin) + conv(f"{name}.conv1", cin, cout, 3)
    if temb:
        s += linear(f"{name}.time_emb_proj", temb, cout)
    s += norm(f"{name}.norm2", cout) + conv(f"{name}.conv2", cout, cout, 3)
    if cin != cout:
        s += conv(f"{name}.conv_shortcut", cin, cout, 1)
    return s


def _transformer_spec(name: str, c: int) -> List[P]:
    tb = f"{name}.transformer_blocks.0"
    s = norm(f"{name}.norm", c) + conv(f"{name}.proj_in", c, c, 1)
    s += norm(f"{tb}.norm1", c)
    s += linear(f"{tb}.attn1.to_q", c, c, False) + linear(f"{tb}.attn1.to_k", c, c, False)
    s += linear(f"{tb}.attn1.to_v", c, c, False) + linear(f"{tb}.attn1.to_out.0", c, c)
    s += norm(f"{tb}.norm2", c)
    s += linear(f"{tb}.attn2.to_q", c, c, False) + linear(f"{tb}.attn2.to_k", CROSS_DIM, c, False)
    s += linear(f"{tb}.attn2.to_v", CROSS_DIM, c, False) + linear(f"{tb}.attn2.to_out.0", c, c)
    s += norm(f"{tb}.norm3", c)
    s += linear(f"{tb}.ff.net.0.proj", c, 8 * c) + linear(f"{tb}.ff.net.2", 4 * c, c)
    s += conv(f"{name}.proj_out", c, c, 1)
    return s


def unet_topology():
    """Static description of the SD-1.5 UNet used by both the spec and the forward pass.

    Returns dict with 'down', 'mid', 'up' lists of (resnet (cin,cout), has_attn) and sampler flags.
    """
    down = []
    cin = BLOCK_OUT[0]
    for i, cout in enumerate(BLOCK_OUT):
        layers = []
        for j in range(2):
            layers.append(((cin if j == 0 else cout), cout))
        down.append(dict(resnets=layers, attn=(i < 3), downsample=(i < 3), c=cout))
        cin = cout
    # skip channel list, in push order
    skips = [BLOCK_OUT[0]]
    for blk in down:
        for _ in blk["resnets"]:
            skips.append(blk["c"])
        if blk["downsample"]:
            skips.append(blk["c"])
    up = []
    rev = list(reversed(BLOCK_OUT))
    prev = rev[0]
    sk = list(skips)
    for i, cout in enumerate(rev):
        layers = []
        for j in range(3):
            s = sk.pop()
            layers.append(((prev if j == 0 else cout) + s, cout))
        up.append(dict(resnets=layers, attn=(i > 0), upsample=(i < 3), c=cout))
        prev = cout
    return dict(down=down, up=up, mid_c=BLOCK_OUT[-1])


def unet_spec() -> List[P]:
    topo = unet_topology()
    s = linear("time_embedding.linear_1", BLOCK_OUT[0], TEMB) + linear("time_embedding.linear_2", TEMB, TEMB)
    s += conv("conv_in", 4, BLOCK_OUT[0], 3)
    for i, blk in enumerate(topo["down"]):
        for j, (cin, cout) in enumerate(blk["resnets"]):
            s += _resnet_spec(f"down_blocks.{i}.resnets.{j}", cin, cout)
            if blk["attn"]:
                s += _transformer_spec(f"down_blocks.{i}.attentions.{j}", cout)
        if blk["downsample"]:
            s += conv(f"down_blocks.{i}.downsamplers.0.conv", blk["c"], blk["c"], 3)
    c = topo["mid_c"]
    s += _resnet_spec("mid_block.resnets.0", c, c) + _transformer_spec("mid_block.attentions.0", c)
    s += _resnet_spec("mid_block.resnets.1", c, c)
    for i, blk in enumerate(topo["up"]):
        for j, (cin, cout) in enumerate(blk["resnets"]):
            s += _resnet_spec(f"up_blocks.{i}.resnets.{j}", cin, cout)
            if blk["attn"]:
                s += _transformer_spec(f"up_blocks.{i}.attentions.{j}", cout)
        if blk["upsample"]:
            s += conv(f"up_blocks.{i}.upsamplers.0.conv", blk["c"], blk["c"], 3)
    s += norm("conv_norm_out", BLOCK_OUT[0]) + conv("conv_out", BLOCK_OUT[0], 4, 3)
    return s


def attn_processor_names() -> List[str]:
    """Keys of diffusers' ``unet.attn_processors`` for SD-1.5."""
    names = []
    topo = unet_topology()
    for i, blk in enumerate(topo["down"]):
        if blk["attn"]:
            for j in range(2):
                for a in ("attn1", "attn2"):
                    names.append(f"down_blocks.{i}.attentions.{j}.transformer_blocks.0.{a}.processor")
    for i, blk in enumerate(topo["up"]):
        if blk["attn"]:
            for j in range(3):
                for a in ("attn1", "attn2"):
                    names.append(f"up_blocks.{i}.attentions.{j}.transformer_blocks.0.{a}.processor")
    for a in ("attn1", "attn2"):
        names.append(f"mid_block.attentions.0.transformer_blocks.0.{a}.processor")
    return names


VAE_BLOCK_OUT = (128, 256, 512, 512)


def vae_decoder_spec() -> List[P]:
    s = conv("post_quant_conv", 4, 4, 1)
    c = VAE_BLOCK_OUT[-1]
    s += conv("decoder.conv_in", 4, c, 3)
    s += _resnet_spec("decoder.mid_block.resnets.0", c, c, None)
    a = "decoder.mid_block.attentions.0"
    s += norm(f"{a}.group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s += linear(f"{a}.{n}", c, c)
    s += _resnet_spec("decoder.mid_block.resnets.1", c, c, None)
    prev = c
    for i, cout in enumerate(reversed(VAE_BLOCK_OUT)):
        for j in range(3):
            s += _resnet_spec(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout, None)
        if i < 3:
            s += conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        prev = cout
    s += norm("decoder.conv_norm_out", VAE_BLOCK_OUT[0]) + conv("decoder.conv_out", VAE_BLOCK_OUT[0], 3, 3)
    return s


def vae_encoder_spec() -> List[P]:
    """AutoencoderKL encoder + quant_conv of SD-1.5 (diffusers==0.23.1 `Encoder`: in 3, block_out (128,256,512,512),
    2 resnets per block, Downsample2D(padding=0) after blocks 0-2, mid block with one single-head attention,
    double_z conv_out 512 -> 8).  34,163,592 + 72 parameters (with the decoder's 49,490,199: the 83,653,863 of the
    published SD-1.5 VAE)."""
    s = conv("encoder.conv_in", 3, VAE_BLOCK_OUT[0], 3)
    prev = VAE_BLOCK_OUT[0]
    for i, cout in enumerate(VAE_BLOCK_OUT):
        for j in range(2):
            s += _resnet_spec(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else cout, cout, None)
        if i < 3:
            s += conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        prev = cout
    c = VAE_BLOCK_OUT[-1]
    s += _resnet_spec("encoder.mid_block.resnets.0", c, c, None)
    a = "encoder.mid_block.attentions.0"
    s += norm(f"{a}.group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        s += linear(f"{a}.{n}", c, c)
    s += _resnet_spec("encoder.mid_block.resnets.1", c, c, None)
    s += norm("encoder.conv_norm_out", c) + conv("encoder.conv_out", c, 8, 3)
    s += conv("quant_conv", 8, 8, 1)
    return s


# =======================================================================================
# Forward pieces
# =======================================================================================
def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """flip_sin_to_cos=True, freq_shift=0: cat[cos, sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)


def _gn(sd, name, x, eps, groups=GROUPS):
    return F.group_norm(x, groups, sd[f"{name}.weight"], sd[f"{name}.bias"], eps)


def _conv(sd, name, x, stride=1, padding=1):
    return F.conv2d(x, sd[f"{name}.weight"], sd[f"{name}.bias"], stride=stride, padding=padding)


def _lin(sd, name, x):
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def resnet_block(sd, name, x, temb_act, eps=1e-5):
    h = _conv(sd, f"{name}.conv1", F.silu(_gn(sd, f"{name}.norm1", x, eps)))
    if temb_act is not None:
        h = h + _lin(sd, f"{name}.time_emb_proj", temb_act)[:, :, None, None]
    h = _conv(sd, f"{name}.conv2", F.silu(_gn(sd, f"{name}.norm2", h, eps)))
    if f"{name}.conv_shortcut.weight" in sd:
        x = _conv(sd, f"{name}.conv_shortcut", x, padding=0)
    return x + h


def mha(q, k, v, heads: int):
    """softmax(q k^T / sqrt(d)) v with [B, N, heads*d] tensors; fp32."""
    B, N, C = q.shape
    d = C // heads
    q = q.view(B, N, heads, d).transpose(1, 2)
    k = k.view(B, k.shape[1], heads, d).transpose(1, 2)
    v = v.view(B, v.shape[1], heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    o = torch.matmul(torch.softmax(s, dim=-1), v)
    return o.transpose(1, 2).reshape(B, N, C)


def plain_attention(sd, name, h, ctx):
    """diffusers ``Attention`` with the default processor (bias-free q/k/v, biased out)."""
    ctx = h if ctx is None else ctx
    o = mha(_lin(sd, f"{name}.to_q", h), _lin(sd, f"{name}.to_k", ctx), _lin(sd, f"{name}.to_v", ctx), HEADS)
    return _lin(sd, f"{name}.to_out.0", o)


Attn2Fn = Callable[[Dict[str, torch.Tensor], str, torch.Tensor, torch.Tensor], torch.Tensor]


def transformer_2d(sd, name, x, ctx, attn2: Optional[Attn2Fn] = None):
    B, C, H, W = x.shape
    res = x
    h = _conv(sd, f"{name}.proj_in", _gn(sd, f"{name}.norm", x, 1e-6), padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    tb = f"{name}.transformer_blocks.0"
    ln = lambda n, z: F.layer_norm(z, (C,), sd[f"{tb}.{n}.weight"], sd[f"{tb}.{n}.bias"], 1e-5)
    h = h + plain_attention(sd, f"{tb}.attn1", ln("norm1", h), None)
    hn = ln("norm2", h)
    if attn2 is None:
        h = h + plain_attention(sd, f"{tb}.attn2", hn, ctx)
    else:
        h = h + attn2(sd, f"{tb}.attn2", hn, ctx)
    hn = ln("norm3", h)
    ag = _lin(sd, f"{tb}.ff.net.0.proj", hn)
    a, g = ag.chunk(2, dim=-1)
    h = h + _lin(sd, f"{tb}.ff.net.2", a * F.gelu(g))
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv(sd, f"{name}.proj_out", h, padding=0) + res


def unet_forward(sd, x, t, ctx, attn2: Optional[Attn2Fn] = None, taps: Optional[dict] = None):
    """x [B,4,H,W] fp32, t [B] (or scalar) timesteps, ctx [B,77,768] -> eps [B,4,H,W].

    ``attn2(sd, attn_name, hidden[B,N,C], ctx)`` replaces the cross-attention (the reference's
    AudioAttnProcessor hook, models/audio_attention_processor.py:43-145).  ``taps`` (optional dict)
    receives named intermediate activations for kernel-level debugging.
    """
    topo = unet_topology()
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.float32, device=x.device)
    t = t.reshape(-1).float().expand(x.shape[0]) if t.numel() == 1 else t.float()
    temb = timestep_embedding(t, BLOCK_OUT[0])
    temb = _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", temb)))
    temb_act = F.silu(temb)
    h = _conv(sd, "conv_in", x)
    if taps is not None:
        taps["conv_in"] = h
    skips = [h]
    for i, blk in enumerate(topo["down"]):
        for j in range(len(blk["resnets"])):
            h = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", h, temb_act)
            if blk["attn"]:
                h = transformer_2d(sd, f"down_blocks.{i}.attentions.{j}", h, ctx, attn2)
            skips.append(h)
        if blk["downsample"]:
            h = _conv(sd, f"down_blocks.{i}.downsamplers.0.conv", h, stride=2)
            skips.append(h)
        if taps is not None:
            taps[f"down{i}"] = h
    h = resnet_block(sd, "mid_block.resnets.0", h, temb_act)
    h = transformer_2d(sd, "mid_block.attentions.0", h, ctx, attn2)
    h = resnet_block(sd, "mid_block.resnets.1", h, temb_act)
    if taps is not None:
        taps["mid"] = h
    for i, blk in enumerate(topo["up"]):
        for j in range(len(blk["resnets"])):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", h, temb_act)
            if blk["attn"]:
                h = transformer_2d(sd, f"up_blocks.{i}.attentions.{j}", h, ctx, attn2)
        if blk["upsample"]:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"up_blocks.{i}.upsamplers.0.conv", h)
        if taps is not None:
            taps[f"up{i}"] = h
    h = F.silu(_gn(sd, "conv_norm_out", h, 1e-5))
    return _conv(sd, "conv_out", h)


# =======================================================================================
# VAE decoder
# =======================================================================================
VAE_SCALING = 0.18215


def vae_decode(sd, z):
    """AutoencoderKL.decode(z / 0.18215) -> [B,3,8H,8W] in roughly [-1,1]."""
    z = z / VAE_SCALING
    h = _conv(sd, "post_quant_conv", z, padding=0)
    h = _conv(sd, "decoder.conv_in", h)
    h = resnet_block(sd, "decoder.mid_block.resnets.0", h, None, eps=1e-6)
    a = "decoder.mid_block.attentions.0"
    B, C, H, W = h.shape
    hn = _gn(sd, f"{a}.group_norm", h, 1e-6).view(B, C, H * W).transpose(1, 2)
    o = mha(_lin(sd, f"{a}.to_q", hn), _lin(sd, f"{a}.to_k", hn), _lin(sd, f"{a}.to_v", hn), 1)
    o = _lin(sd, f"{a}.to_out.0", o).transpose(1, 2).reshape(B, C, H, W)
    h = h + o
    h = resnet_block(sd, "decoder.mid_block.resnets.1", h, None, eps=1e-6)
    for i in range(4):
        for j in range(3):
            h = resnet_block(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, None, eps=1e-6)
        if i < 3:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", h)
    h = F.silu(_gn(sd, "decoder.conv_norm_out", h, 1e-6))
    return _conv(sd, "decoder.conv_out", h)


def vae_encode(sd, img):
    """AutoencoderKL.encode(img).latent_dist -> (mean, logvar), each [B,4,H/8,W/8]; the latent the pipeline stores is
    `mean * 0.18215` (mode of the posterior; data/audiocaps_latent_v4.py:185 keeps (4,64,64) latents)."""
    h = _conv(sd, "encoder.conv_in", img)
    for i in range(4):
        for j in range(2):
            h = resnet_block(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, None, eps=1e-6)
        if i < 3:
            h = F.pad(h, (0, 1, 0, 1))                                   # Downsample2D(padding=0): right / bottom only
            h = _conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", h, stride=2, padding=0)
    h = resnet_block(sd, "encoder.mid_block.resnets.0", h, None, eps=1e-6)
    a = "encoder.mid_block.attentions.0"
    B, C, H, W = h.shape
    hn = _gn(sd, f"{a}.group_norm", h, 1e-6).view(B, C, H * W).transpose(1, 2)
    o = mha(_lin(sd, f"{a}.to_q", hn), _lin(sd, f"{a}.to_k", hn), _lin(sd, f"{a}.to_v", hn), 1)
    o = _lin(sd, f"{a}.to_out.0", o).transpose(1, 2).reshape(B, C, H, W)
    h = h + o
    h = resnet_block(sd, "encoder.mid_block.resnets.1", h, None, eps=1e-6)
    h = F.silu(_gn(sd, "encoder.conv_norm_out", h, 1e-6))
    m = _conv(sd, "quant_conv", _conv(sd, "encoder.conv_out", h), padding=0)
    mean, logvar = m[:, :4], m[:, 4:].clamp(-30.0, 20.0)
    return mean, logvar


# =======================================================================================
# Schedulers
# =======================================================================================
def alphas_cumprod() -> torch.Tensor:
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(n: int) -> List[int]:
    ratio = 1000 // n
    return [(n - 1 - i) * ratio + 1 for i in range(n)]


def ddim_coeffs(n: int) -> List[Tuple[int, float, float]]:
    """Per step (t, a, b) with x_next = a*x + b*eps (eta=0, eps-prediction, steps_offset=1,
    set_alpha_to_one=False)."""
    ac = alphas_cumprod()
    ratio = 1000 // n
    out = []
    for t in ddim_timesteps(n):
        tp = t - ratio
        a_t = float(ac[t])
        a_p = float(ac[tp]) if tp >= 0 else float(ac[0])
        # x0 = (x - sqrt(1-a_t) eps)/sqrt(a_t); x' = sqrt(a_p) x0 + sqrt(1-a_p) eps
        ca = math.sqrt(a_p / a_t)
        cb = math.sqrt(1.0 - a_p) - math.sqrt(a_p) * math.sqrt(1.0 - a_t) / math.sqrt(a_t)
        out.append((t, ca, cb))
    return out


def euler_sigmas(n: int) -> Tuple[List[float], List[float]]:
    """'leading' spacing with steps_offset=1.  Returns (timesteps, sigmas[n+1])."""
    import numpy as np
    ac = alphas_cumprod().numpy()
    sig_all = ((1 - ac) / ac) ** 0.5
    ts = np.array(ddim_timesteps(n), dtype=np.float64)
    sig = np.interp(ts, np.arange(1000), sig_all)
    return [float(v) for v in ts], [float(v) for v in sig] + [0.0]


def cfg_combine(eps2: torch.Tensor, g: float) -> torch.Tensor:
    """eps2 = cat[uncond, cond] on batch."""
    eu, ec = eps2.chunk(2, dim=0)
    return eu + g * (ec - eu)
