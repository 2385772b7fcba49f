"""fp32 CPU restatement of the intended denoising pipeline (decisions D1-D4).
TEST INFRASTRUCTURE (see oracle/__init__.py).

clap[B,512] -> ImprovedHierarchicalAudioEncoder -> routed{early,mid,late} -> AudioAttnProcessor
(mode 'add' | 'concat') on the 16 attn2 sites of the SD-1.5 UNet; DDIM (eta 0) or Euler; CFG with
cat[uncond, cond]; VAE decode.  The reference has no call site for this loop
(scripts/inference.py:153-166 fabricates a random image), so the wiring is ours (D1).
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Optional

import numpy as np
import torch

from . import audio as A
from . import sd15
from .weights import synth_state_dict

LEVELS = ("early", "mid", "late")


# ---------------------------------------------------------------------------------------
# synthetic inputs (portable: numpy PCG64)
# ---------------------------------------------------------------------------------------
def text_states(prompt: str) -> np.ndarray:
    """D4: synthetic CLIP-like text states [77,768], keyed by a hash of the prompt."""
    g = np.random.Generator(np.random.PCG64(zlib.crc32(prompt.encode()) + 0x7E57))
    return g.standard_normal((77, 768), dtype=np.float32)


def init_noise(seed: int, h: int = 64, w: int = 64) -> np.ndarray:
    g = np.random.Generator(np.random.PCG64(int(seed) + 0x4015E))
    return g.standard_normal((4, h, w), dtype=np.float32)


def np_randn(tag: str, shape, seed: int = 0) -> np.ndarray:
    """Portable named standard-normal test tensor."""
    g = np.random.Generator(np.random.PCG64((zlib.crc32(tag.encode()) << 20) + int(seed)))
    return g.standard_normal(tuple(shape), dtype=np.float32)


def clap_embedding(seed: int) -> np.ndarray:
    """Stand-in for CLAPAudioEncoder.encode_audio output: unit-norm [512] (audio_encoder.py:171-174;
    scripts/inference.py:85-90 itself uses randn(1,512))."""
    g = np.random.Generator(np.random.PCG64(int(seed) + 0xC1A9))
    e = g.standard_normal((512,), dtype=np.float32)
    return e / np.linalg.norm(e)


def synthetic_audio(seed: int, n: int = 480000) -> np.ndarray:
    """10 s @ 48 kHz mono 0.1*randn, peak-normalised as scripts/inference.py:81."""
    g = np.random.Generator(np.random.PCG64(int(seed) + 0xA0D10))
    a = 0.1 * g.standard_normal((n,), dtype=np.float32)
    return a / (np.abs(a).max() + 1e-8)


CONFIG3_PROMPTS = ("a beach", "a city street", "a forest", "a thunderstorm", "a cafe", "a train", "a river", "a crowd")


def config3_jobs(n: int = 8, first_seed: int = 100):
    """(prompt, seed) jobs of one micro-batch of the config-3 sweep (BASELINE.json configs[2]: 8 prompts x 8 seeds)."""
    return [(CONFIG3_PROMPTS[j % len(CONFIG3_PROMPTS)], first_seed + j) for j in range(n)]


# ---------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------
def to_torch(sd: Dict[str, np.ndarray], device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(v).to(device=device, dtype=dtype) for k, v in sd.items()}


def build_weights(seed: int = 0, with_vae: bool = True, device="cpu") -> Dict[str, Dict[str, torch.Tensor]]:
    """All synthetic weights of the pipeline, as torch fp32 dicts."""
    out = {"unet": to_torch(synth_state_dict(sd15.unet_spec(), seed), device)}
    hier = synth_state_dict(A.improved_hier_spec(), seed)
    for k, v in A.IMPROVED_BUFFERS.items():
        hier[k] = np.asarray(v, dtype=np.float32)
    out["hier"] = to_torch(hier, device)
    for lvl in LEVELS:
        out[f"proc_{lvl}"] = to_torch(synth_state_dict(A.attn_processor_spec(), seed, prefix=f"{lvl}."), device)
        out[f"proc_{lvl}"] = {k[len(lvl) + 1:]: v for k, v in out[f"proc_{lvl}"].items()}
    out["adapter"] = to_torch(synth_state_dict(A.audio_adapter_spec(), seed), device)
    if with_vae:
        out["vae"] = to_torch(synth_state_dict(sd15.vae_decoder_spec(), seed), device)
    return out


# ---------------------------------------------------------------------------------------
# the loop
# ---------------------------------------------------------------------------------------
def make_attn2_hook(weights, routed: Dict[str, torch.Tensor], mode: str = "add"):
    """Returns attn2(sd, name, h, ctx) implementing AudioAttnProcessor at every attn2 site
    (level by AudioProcessorManager's name rules, one shared processor per level)."""
    def hook(sd, name, h, ctx):
        lvl = A.level_of_site(name)
        attn = {k: sd[f"{name}.{k}"] for k in ("to_q.weight", "to_k.weight", "to_v.weight",
                                                "to_out.0.weight", "to_out.0.bias")}
        return A.processor_call(weights[f"proc_{lvl}"], attn, sd15.HEADS, h, ctx, routed[lvl], mode)
    return hook


@torch.no_grad()
def sample(weights, clap: torch.Tensor, ctx_cond: torch.Tensor, ctx_uncond: torch.Tensor,
           noise: torch.Tensor, steps: int = 50, guidance: float = 7.5, scheduler: str = "ddim",
           mode: str = "add", use_audio: bool = True, max_steps: Optional[int] = None,
           decode: bool = False) -> Dict[str, object]:
    """clap [B,512], ctx_* [B,77,768], noise [B,4,H,W] (all fp32, same device).

    Returns dict(latents=[per-step latents], eps=[per-step guided eps], image (if decode))."""
    B = noise.shape[0]
    hier = A.improved_hier_forward(weights["hier"], clap)
    routed = {k: torch.cat([v, v], dim=0) for k, v in hier["routed"].items()}     # D2: same audio on both halves
    hook = make_attn2_hook(weights, routed, mode) if use_audio else None
    ctx2 = torch.cat([ctx_uncond, ctx_cond], dim=0)
    x = noise.clone()
    lat, epss = [], []
    if scheduler == "ddim":
        plan = sd15.ddim_coeffs(steps)
        for i, (t, ca, cb) in enumerate(plan):
            if max_steps is not None and i >= max_steps:
                break
            eps2 = sd15.unet_forward(weights["unet"], torch.cat([x, x], 0), float(t), ctx2, hook)
            eps = sd15.cfg_combine(eps2, guidance)
            x = ca * x + cb * eps
            lat.append(x.clone()); epss.append(eps)
    elif scheduler == "euler":
        ts, sig = sd15.euler_sigmas(steps)
        x = x * sig[0]
        for i, t in enumerate(ts):
            if max_steps is not None and i >= max_steps:
                break
            xin = x / (sig[i] ** 2 + 1.0) ** 0.5
            eps2 = sd15.unet_forward(weights["unet"], torch.cat([xin, xin], 0), float(t), ctx2, hook)
            eps = sd15.cfg_combine(eps2, guidance)
            x = x + eps * (sig[i + 1] - sig[i])
            lat.append(x.clone()); epss.append(eps)
    else:
        raise ValueError(scheduler)
    out = dict(latents=lat, eps=epss, routed=hier["routed"], tokens_77=hier["tokens_77"])
    if decode:
        out["image"] = sd15.vae_decode(weights["vae"], x)
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 2.0) -> float:
    """Images in [-1,1] -> peak-to-peak 2."""
    mse = float(((a.double().cpu() - b.double().cpu()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * float(np.log10(peak * peak / mse))
