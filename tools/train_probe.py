#!/usr/bin/env python
"""Stage-3 step probe (GPU): one forward_backward in the given dtype at a small latent; CUDA_LAUNCH_BLOCKING=1 localises faults."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from clap2diffusion_b200 import ops, synthetic  # noqa: E402
from clap2diffusion_b200 import unet as unet_mod  # noqa: E402
from clap2diffusion_b200.models.hierarchical_audio_v4 import ImprovedHierarchicalAudioEncoder  # noqa: E402
from clap2diffusion_b200.train import Stage3Trainer  # noqa: E402

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
usd = synthetic.random_state_dict(unet_mod.param_shapes(), 0, "cuda")
hier = ImprovedHierarchicalAudioEncoder().to("cuda").eval()
tr = Stage3Trainer(usd, hier, None, device="cuda", dtype=dt)
g = torch.Generator().manual_seed(1234)
batch = {"audio_embedding": torch.from_numpy(np.stack([synthetic.clap_embedding(k) for k in range(B)])),
         "image_latents": torch.randn(B, 4, hw, hw, generator=g),
         "text_embedding": torch.from_numpy(np.stack([synthetic.text_states(f"prompt {k % 8}") for k in range(B)])),
         "noise": torch.randn(B, 4, hw, hw, generator=g), "timesteps": torch.randint(0, 1000, (B,), generator=g)}
batch = {k: v.to("cuda") for k, v in batch.items()}
torch.cuda.synchronize()
for i in range(3):
    t0 = time.time()
    out = tr.train_step(batch)
    torch.cuda.synchronize()
    print(f"step {i}: loss {float(out['diffusion']):.6f} grad_norm {float(out['grad_norm']):.4e}  {1e3 * (time.time() - t0):.1f} ms")
if os.environ.get("C2D_PROFILE"):            # one step between cudaProfilerStart / Stop (ncu --profile-from-start off)
    torch.cuda.profiler.start()
    tr.train_step(batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
ops.PROFILE = []
tr.train_step(batch)
torch.cuda.synchronize()
rec, ops.PROFILE = ops.PROFILE, None
agg = {}
for name, fl, nb, e0, e1 in rec:
    a = agg.setdefault(name, [0.0, 0])
    a[0] += e0.elapsed_time(e1); a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"timed ops total {tot:.2f} ms")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {k:20s} {a[0]:8.3f} ms x{a[1]}")

# host-side enqueue time vs device time of N steps (is the step launch-bound?)
torch.cuda.synchronize()
N = 5
t0 = time.time()
for _ in range(N):
    tr.train_step(batch)
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
print(f"{N} steps: host enqueue {1e3 * (t1 - t0) / N:.1f} ms/step, until the device is idle {1e3 * (t2 - t0) / N:.1f} ms/step")
