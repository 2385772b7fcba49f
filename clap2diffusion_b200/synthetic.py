"""Synthetic inputs and random-init weights for the product (there is no network for checkpoints,
datasets or tokenizer files; SURVEY.md §8d / decision D4).

Inputs are portable functions of (tag, seed) drawn from numpy's PCG64, so every rank of a data-parallel
run -- and the CPU oracle in the tests -- sees the same prompt states / noise / CLAP stand-in without
communication.  Weights for the bench are drawn directly on the GPU (fan-in-scaled uniform, the bound of
torch's default Linear/Conv init).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import numpy as np
import torch


def text_states(prompt: str) -> np.ndarray:
    """Synthetic CLIP-like text states [77,768] keyed by a hash of the prompt (no tokenizer on the box)."""
    g = np.random.Generator(np.random.PCG64(zlib.crc32(prompt.encode()) + 0x7E57))
    return g.standard_normal((77, 768), dtype=np.float32)


def init_noise(seed: int, h: int = 64, w: int = 64) -> np.ndarray:
    """Initial latent noise [4,h,w], generated on the host so CPU and GPU runs share it."""
    g = np.random.Generator(np.random.PCG64(int(seed) + 0x4015E))
    return g.standard_normal((4, h, w), dtype=np.float32)


def clap_embedding(seed: int) -> np.ndarray:
    """Unit-norm [512] stand-in for CLAPAudioEncoder.encode_audio (reference scripts/inference.py:85-90 uses
    randn(1,512) itself)."""
    g = np.random.Generator(np.random.PCG64(int(seed) + 0xC1A9))
    e = g.standard_normal((512,), dtype=np.float32)
    return e / np.linalg.norm(e)


def synthetic_audio(seed: int, n: int = 480000) -> np.ndarray:
    """10 s @ 48 kHz mono, 0.1*randn, peak-normalised like reference scripts/inference.py:81."""
    g = np.random.Generator(np.random.PCG64(int(seed) + 0xA0D10))
    a = 0.1 * g.standard_normal((n,), dtype=np.float32)
    return a / (np.abs(a).max() + 1e-8)


def random_state_dict(shapes: Dict[str, tuple], seed: int, device="cuda") -> Dict[str, torch.Tensor]:
    """Random-init fp32 weights for a {name: shape} spec, drawn on ``device``:
    matrices / conv kernels U(+-1/sqrt(fan_in)); norm gains 1 + 0.1 U; other vectors 0.05 U."""
    gen = torch.Generator(device=device)
    out = {}
    for i, (name, shape) in enumerate(shapes.items()):
        gen.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        u = torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 2.0 - 1.0
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            u *= 1.0 / math.sqrt(fan_in)
        elif "norm" in name and name.endswith(".weight"):
            u = 1.0 + 0.1 * u
        else:
            u *= 0.05
        out[name] = u
    return out
