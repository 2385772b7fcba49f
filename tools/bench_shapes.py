#!/usr/bin/env python
"""Per-shape CUDA-event timing of the UNet's dense kernels at the benchmark batch (iteration aid, GPU only).

    python tools/bench_shapes.py [--what linear,conv,geglu,attn,gn,ln] [--batch 16] [--reps 20]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from clap2diffusion_b200 import ops  # noqa: E402


GRAPH = True


def timeit(fn, reps):
    """Average device time of one call in us.  By default the `reps` calls are captured into one CUDA graph and the graph is
    replayed, so the Python / ctypes / tensor-map-encode cost of a call (~30 us, more than most of these kernels) is not in
    the number; --no-graph times eager launches."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if GRAPH:
        st = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(st):
            fn()
            with torch.cuda.graph(g, stream=st):
                for _ in range(reps):
                    fn()
            g.replay()
            st.synchronize()
            e0.record(st)
            g.replay()
            e1.record(st)
            st.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="linear,conv,geglu,attn,gn,ln")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-graph", action="store_true", help="eager launches (includes the host-side cost of a call)")
    a = ap.parse_args()
    global GRAPH
    GRAPH = not a.no_graph
    what = set(a.what.split(","))
    dev = torch.device("cuda", 0)
    B = a.batch
    bf = torch.bfloat16
    g = torch.Generator(device=dev).manual_seed(0)

    def rn(*s):
        return (torch.randn(*s, device=dev, generator=g) * 0.5).to(bf)

    levels = [(64, 320), (32, 640), (16, 1280), (8, 1280)]
    if "linear" in what:
        print("linear: M N K residual  us  TF/s  GB/s")
        for hw, C in levels:
            M = B * hw * hw
            for (N, K, res) in [(C, C, False), (C, C, True), (3 * C, C, False), (C, 4 * C, True), (C, 2 * C, False), (C, 3 * C, False)]:
                x, w = rn(M, K), rn(N, K)
                bias = torch.zeros(N, device=dev)
                r = rn(M, N) if res else None
                out = torch.empty(M, N, device=dev, dtype=bf)
                us = timeit(lambda: ops.linear(x, w, bias, residual=r, out=out), a.reps)
                fl = 2.0 * M * N * K
                by = 2.0 * (M * K + N * K + M * N * (2 if res else 1))
                print(f"  {M:6d} {N:5d} {K:5d} {int(res)}  {us:8.1f}  {fl / us / 1e6:7.1f}  {by / us / 1e3:7.1f}")
    if "geglu" in what:
        print("geglu: M F K  us  TF/s")
        for hw, C in levels:
            M = B * hw * hw
            F, K = 4 * C, C
            x = rn(M, K)
            w = torch.randn(2 * F, K, device=dev, generator=g) * 0.05
            bias = torch.zeros(2 * F, device=dev)
            wp, bp = ops.pack_geglu(w, bias, bf)
            out = torch.empty(M, F, device=dev, dtype=bf)
            us = timeit(lambda: ops.geglu_linear(x, wp, bp, out=out), a.reps)
            print(f"  {M:6d} {F:5d} {K:5d}  {us:8.1f}  {4.0 * M * F * K / us / 1e6:7.1f}")
    if "conv" in what:
        print("conv3x3: H Cin Cout stride  us  TF/s")
        for (hw, cin, cout, st) in [(64, 320, 320, 1), (64, 640, 320, 1), (64, 960, 320, 1), (64, 320, 320, 2), (32, 320, 640, 1),
                                    (32, 640, 640, 1), (32, 1280, 640, 1), (32, 1920, 640, 1), (32, 960, 640, 1), (32, 640, 640, 2),
                                    (16, 640, 1280, 1), (16, 1280, 1280, 1), (16, 2560, 1280, 1), (16, 1920, 1280, 1), (16, 1280, 1280, 2),
                                    (8, 1280, 1280, 1), (8, 2560, 1280, 1)]:
            x = rn(B, hw, hw, cin)
            w = ops.pack_conv3x3(torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.02, bf)
            bias = torch.zeros(cout, device=dev)
            ho = hw // st
            out = torch.empty(B, ho, ho, cout, device=dev, dtype=bf)
            us = timeit(lambda: ops.conv3x3(x, w, bias, stride=st, out=out), a.reps)
            print(f"  {hw:3d} {cin:5d} {cout:5d} {st}  {us:8.1f}  {2.0 * B * ho * ho * cout * 9 * cin / us / 1e6:7.1f}")
    if "attn" in what:
        print("attention: N Nkv C d  us  TF/s(alg)")
        for hw, C in levels:
            N = hw * hw
            d = C // 8
            qkv = rn(B, N, 3 * C)
            out = torch.empty(B, N, C, device=dev, dtype=bf)
            us = timeit(lambda: ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 8, out=out), a.reps)
            print(f"  {N:5d} {N:5d} {C:5d} {d:4d}  {us:8.1f}  {4.0 * B * 8 * N * N * d / us / 1e6:7.1f}")
            q, kv = rn(B, N, C), rn(B, 77, 2 * C)
            us = timeit(lambda: ops.attention(q, kv[..., :C], kv[..., C:], 8, out=out), a.reps)
            print(f"  {N:5d} {77:5d} {C:5d} {d:4d}  {us:8.1f}  {4.0 * B * 8 * N * 77 * d / us / 1e6:7.1f}")
    if "gn" in what:
        print("group_norm(+silu): HW C1 C2  us  GB/s (alg: read x2 + write)")
        for (hw, c1, c2) in [(64, 320, 0), (64, 320, 320), (64, 640, 320), (32, 640, 0), (32, 320, 0), (32, 640, 640), (32, 1280, 640),
                             (32, 640, 320), (16, 1280, 0), (16, 640, 0), (16, 1280, 1280), (16, 1280, 640), (8, 1280, 0), (8, 1280, 1280)]:
            x = rn(B, hw * hw, c1)
            x2 = rn(B, hw * hw, c2) if c2 else None
            C = c1 + c2
            gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            out = torch.empty(B, hw * hw, C, device=dev, dtype=bf)
            us = timeit(lambda: ops.group_norm(x, gam, bet, 32, 1e-5, True, x2=x2, out=out), a.reps)
            s1 = ops.channel_stats(x, torch.zeros(B * c1 * 2, device=dev, dtype=torch.int64))
            s2 = ops.channel_stats(x2, torch.zeros(B * c2 * 2, device=dev, dtype=torch.int64)) if c2 else None
            us2 = timeit(lambda: ops.group_norm_apply(x, s1, gam, bet, 32, 1e-5, True, x2=x2, stats2=s2, out=out), a.reps)
            print(f"  {hw * hw:5d} {c1:5d} {c2:5d}  {us:8.1f}  {3.0 * B * hw * hw * C * 2 / us / 1e3:7.1f}   one-pass apply {us2:8.1f} us "
                  f"{2.0 * B * hw * hw * C * 2 / us2 / 1e3:7.1f} GB/s")
    if "ln" in what:
        print("layer_norm: M C  us  GB/s")
        for hw, C in levels:
            M = B * hw * hw
            x = rn(M, C)
            gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
            out = torch.empty_like(x)
            us = timeit(lambda: ops.layer_norm(x, gam, bet, out=out), a.reps)
            print(f"  {M:6d} {C:5d}  {us:8.1f}  {2.0 * M * C * 2 / us / 1e3:7.1f}")


if __name__ == "__main__":
    main()
