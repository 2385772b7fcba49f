#!/usr/bin/env python
"""Audio-side modules (projector / decomposer MLPs) in bf16 on the tcgen05 GEMM: error against the reference goldens and
the kernels each op resolves to (GPU only).  Lives under tests/ because it checks against the oracle (test infrastructure)."""
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import audio as A  # noqa: E402
from oracle.pipeline import rel_l2, to_torch  # noqa: E402
from oracle.weights import synth_state_dict  # noqa: E402
from clap2diffusion_b200 import ops  # noqa: E402
from clap2diffusion_b200.models import audio_adapter_v4 as padapter  # noqa: E402
from clap2diffusion_b200.models import hierarchical_audio_v4 as phier  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DEV = "cuda"


def kernels(fn):
    ops.PROFILE = []
    out = fn()
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    return out, Counter(r[0] for r in rec)


def main():
    g = np.load(os.path.join(GOLD, "improved_hier.npz"))
    sd = synth_state_dict(A.improved_hier_spec(), int(g["seed"]))
    for k, v in A.IMPROVED_BUFFERS.items():
        sd[k] = np.asarray(v, dtype=np.float32)
    m = phier.ImprovedHierarchicalAudioEncoder()
    m.load_state_dict(to_torch(sd))
    m = m.to(DEV).eval()
    clap = torch.from_numpy(g["clap"]).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        with torch.no_grad():
            enc, ks = kernels(lambda: m.encode(clap.to(dt), with_tokens77=True))
        errs = {k: rel_l2(enc[k].float(), torch.from_numpy(g[k])) for k in ("tokens_10", "assignments", "hierarchy_weights", "tokens_77")}
        errs.update({f"routed_{l}": rel_l2(enc["routed"][l].float(), torch.from_numpy(g[f"routed_{l}"])) for l in ("early", "mid", "late")})
        print(dt, {k: f"{v:.1e}" for k, v in errs.items()})
        print("   kernels:", dict(ks))
    ga = np.load(os.path.join(GOLD, "audio_adapter.npz"))
    ad = padapter.AudioAdapter()
    ad.load_state_dict(to_torch(synth_state_dict(A.audio_adapter_spec(), int(ga["seed"]))))
    ad = ad.to(DEV).eval()
    ca = torch.from_numpy(ga["clap"]).to(DEV)
    for dt in (torch.float32, torch.bfloat16):
        with torch.no_grad():
            out, ks = kernels(lambda: ad(ca.to(dt)))
        print(dt, "adapter tokens", f"{rel_l2(out.float(), torch.from_numpy(ga['tokens'])):.1e}", dict(ks))
    # timing at the benchmark micro-batch (8 images) and at config-4 batch 256
    for B in (8, 256):
        x = torch.randn(B, 512, device=DEV)
        x = x / x.norm(dim=-1, keepdim=True)
        for dt in (torch.float32, torch.bfloat16):
            xx = x.to(dt)
            with torch.no_grad():
                for _ in range(3):
                    m.encode(xx, with_tokens77=False); ad(xx)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    m.encode(xx, with_tokens77=False)
                e1.record(); torch.cuda.synchronize()
                t_enc = e0.elapsed_time(e1) / 10
                e0.record()
                for _ in range(10):
                    ad(xx)
                e1.record(); torch.cuda.synchronize()
                print(f"batch {B} {dt}: hier.encode {t_enc * 1e3:.0f} us, AudioAdapter {e0.elapsed_time(e1) / 10 * 1e3:.0f} us")


if __name__ == "__main__":
    main()
