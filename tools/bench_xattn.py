#!/usr/bin/env python
"""Cross-attention site at the benchmark batch: fused c2d_xattn_fwd vs the three-kernel composition (iteration aid
and ncu target, GPU only).

    python tools/bench_xattn.py [--batch 16] [--reps 20] [--only 0] [--fused-only]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from clap2diffusion_b200 import ops  # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--fused-only", action="store_true")
    ap.add_argument("--T", type=int, default=77)
    ap.add_argument("--T2", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    bf, B, heads = torch.bfloat16, a.batch, 8
    g = torch.Generator(device=dev).manual_seed(0)

    def rn(*s, sc=0.5):
        return (torch.randn(*s, device=dev, generator=g) * sc).to(bf)

    print("site: Nq C | fused us  TF/s(alg) | to_q + attn us (3-kernel path, without to_out) | max rel diff")
    for i, (hw, C) in enumerate([(64, 320), (32, 640), (16, 1280), (8, 1280)]):
        if a.only >= 0 and i != a.only:
            continue
        Nq, M = hw * hw, B * hw * hw
        x = rn(B, Nq, C, sc=1.0)
        kv = rn(B, a.T, 2 * C)
        kv2 = rn(B, a.T2, 2 * C) if a.T2 else None
        w32 = torch.randn(C, C, device=dev, generator=g) * C ** -0.5
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ln = ops.pack_lnfold(w32, gamma, beta, None, bf)
        rs = torch.zeros(M * 2, device=dev, dtype=torch.int64)
        ops.linear(x, torch.eye(C, device=dev, dtype=bf), row_stats=rs)
        kvp = ops.xattn_pack_kv(kv, heads, kv2)
        out = torch.empty(B, Nq, C, device=dev, dtype=bf)
        q = torch.empty(B, Nq, C, device=dev, dtype=bf)
        o2 = torch.empty(B, Nq, C, device=dev, dtype=bf)

        def fused():
            ops.xattn(x, kvp, ln=ln, ln_stats=rs, out=out)

        def three():
            ops.linear(x, None, ln=ln, ln_stats=rs, out=q)
            ops.attention(q, kv[..., :C], kv[..., C:], heads, out=o2)

        us_f = timeit(fused, a.reps)
        fl = 2.0 * M * C * C + 4.0 * M * (a.T + a.T2) * C
        if a.fused_only:
            print(f"{Nq:5d} {C:5d} | {us_f:8.1f} {fl / us_f * 1e-6:7.1f}")
            continue
        us_3 = timeit(three, a.reps)
        diff = float((out.float() - o2.float()).norm() / o2.float().norm()) if not a.T2 else float("nan")
        print(f"{Nq:5d} {C:5d} | {us_f:8.1f} {fl / us_f * 1e-6:7.1f} | {us_3:8.1f} | {diff:.2e}")


if __name__ == "__main__":
    main()
