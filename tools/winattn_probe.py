#!/usr/bin/env python
"""Swin window attention at the four HTSAT stages, 256 clips (CUDA-graph timed; C2D_WINATTN=0 selects the first kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from bench_shapes import timeit  # noqa: E402
from clap2diffusion_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for H, C, heads in ((64, 96, 4), (32, 192, 8), (16, 384, 16), (8, 768, 32)):
    for shift in (0, 4):
        if shift and H == 8:
            continue
        qkv = (torch.randn(B, H * H, 3 * C, device="cuda") * 0.5).to(torch.bfloat16)
        bias = torch.randn(heads, 64, 64, device="cuda") * 0.5
        out = torch.empty(B, H * H, C, device="cuda", dtype=torch.bfloat16)
        us = timeit(lambda: ops.window_attention(qkv, bias, H, H, heads, shift, out=out), 5)
        fl = 4.0 * B * H * H * 64 * C
        by = qkv.numel() * 2 + out.numel() * 2
        print(f"H={H:3d} C={C:4d} heads={heads:2d} shift={shift}: {us:9.1f} us  {fl / us / 1e6:6.1f} TF/s  {by / us / 1e3:7.1f} GB/s")
