#!/usr/bin/env python
"""The HTSAT tower's dense layers at 256 clips, one by one (CUDA-graph timed): us, TFLOP/s and GB/s of algorithmic bytes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from bench_shapes import timeit  # noqa: E402
from clap2diffusion_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
bf = torch.bfloat16
dev = "cuda"
tot = 0.0
print("layer                M       N     K  res act      us    TF/s    GB/s   x count")
def run(name, M, N, K, res, act, count):
    global tot
    x = (torch.randn(M, K, device=dev) * 0.5).to(bf)
    w = (torch.randn(N, K, device=dev) * 0.05).to(bf)
    b = torch.zeros(N, device=dev)
    r = (torch.randn(M, N, device=dev) * 0.5).to(bf) if res else None
    out = torch.empty(M, N, device=dev, dtype=bf)
    us = timeit(lambda: ops.linear(x, w, b, residual=r, act=act, out=out), 5)
    by = 2.0 * (M * K + N * K + M * N * (2 if res else 1))
    tot += us * count
    print(f"{name:16s} {M:8d} {N:5d} {K:5d}  {int(res)}   {act}  {us:8.1f} {2.0 * M * N * K / us / 1e6:7.1f} {by / us / 1e3:7.1f}   x{count}")
run("patch_embed", B * 4096, 96, 16, False, ops.ACT_NONE, 1)
n = B * 4096
for i, depth in enumerate((2, 2, 6, 2)):
    c = 96 * 2 ** i
    run(f"s{i}.qkv", n, 3 * c, c, False, ops.ACT_NONE, depth)
    run(f"s{i}.proj", n, c, c, True, ops.ACT_NONE, depth)
    run(f"s{i}.fc1", n, 4 * c, c, False, ops.ACT_GELU, depth)
    run(f"s{i}.fc1 (no act)", n, 4 * c, c, False, ops.ACT_NONE, 0)
    run(f"s{i}.fc2", n, c, 4 * c, True, ops.ACT_NONE, depth)
    if i < 3:
        run(f"s{i}.merge", n // 4, 2 * c, 4 * c, False, ops.ACT_NONE, 1)
        n //= 4
print(f"sum over the tower: {tot / 1e3:.2f} ms")
for C, rows in ((96, B * 4096), (192, B * 1024), (384, B * 256), (768, B * 64)):
    x = (torch.randn(rows, C, device=dev) * 0.5).to(bf)
    g, bb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty_like(x)
    us = timeit(lambda: ops.layer_norm(x, g, bb, out=out), 5)
    print(f"layer_norm rows={rows:8d} C={C:4d}: {us:8.1f} us  {4.0 * rows * C / us / 1e3:7.1f} GB/s")
