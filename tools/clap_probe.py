#!/usr/bin/env python
"""Per-op CUDA-event breakdown of BASELINE config 4 (CLAP HTSAT tower + hierarchical decomposer) on one GPU.

    python tools/clap_probe.py [clips]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from clap2diffusion_b200 import ops, synthetic  # noqa: E402
from clap2diffusion_b200.models.audio_encoder import CLAPAudioEncoder  # noqa: E402
from clap2diffusion_b200.models.hierarchical_audio_v4 import ImprovedHierarchicalAudioEncoder  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    enc = CLAPAudioEncoder.random_init(seed=0, device="cuda:0", dtype=torch.bfloat16)
    hier = ImprovedHierarchicalAudioEncoder().to(dev).eval()
    base = np.stack([synthetic.synthetic_audio(i) for i in range(8)])
    waves = torch.from_numpy(np.concatenate([base] * ((n + 7) // 8))[:n]).to(dev)

    def step():
        return hier.encode(ops.cast(enc.encode_audio(waves), torch.bfloat16), with_tokens77=True)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{n} clips: {e0.elapsed_time(e1) / 3:.2f} ms per pass")
    ops.PROFILE = []
    step()
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for name, fl, nb, t0, t1 in rec:
        a = agg.setdefault(name, [0.0, 0, 0.0])
        a[0] += t0.elapsed_time(t1)
        a[1] += 1
        a[2] += fl
    tot = sum(a[0] for a in agg.values())
    print(f"timed ops total {tot:.2f} ms")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:20]:
        tf = f"{a[2] / (a[0] * 1e-3) / 1e12:7.1f} TF/s" if a[2] else ""
        print(f"  {k:24s} {a[0]:8.3f} ms x{a[1]:<4d} {tf}")


if __name__ == "__main__":
    main()
