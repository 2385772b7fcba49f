#!/usr/bin/env python
"""BASELINE config 2 in isolation: ONE image (UNet batch 2 with CFG), 50 DDIM steps + VAE decode; ms per image.
Env switches (C2D_SPLITK, C2D_SMALL_BN, C2D_XATTN_P, ...) select kernel variants for A/B runs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import contextlib  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

from clap2diffusion_b200 import synthetic  # noqa: E402
from clap2diffusion_b200.pipeline import AudioToImagePipeline  # noqa: E402

dev = torch.device("cuda", 0)
with contextlib.redirect_stdout(sys.stderr):
    pipe = AudioToImagePipeline.random_init(seed=0, device=dev, dtype=torch.bfloat16)
clap = torch.from_numpy(synthetic.clap_embedding(0)[None]).to(dev)
cond = torch.from_numpy(synthetic.text_states("a beach")[None]).to(dev)
unc = torch.from_numpy(synthetic.text_states("")[None]).to(dev)
noise = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(0)).to(dev)
for _ in range(2):
    pipe.sampler.sample(clap, cond, unc, noise, steps=50, guidance=7.5, decode=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pipe.sampler.sample(clap, cond, unc, noise, steps=50, guidance=7.5, decode=True)
e1.record()
torch.cuda.synchronize()
print(f"config 2: {e0.elapsed_time(e1) / 5:.2f} ms per image  ({' '.join(k + '=' + v for k, v in os.environ.items() if k.startswith('C2D_'))})")
