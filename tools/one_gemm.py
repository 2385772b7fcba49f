#!/usr/bin/env python
"""One bf16 linear, a few launches (ncu target):  python tools/one_gemm.py M N K [act] [residual]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from clap2diffusion_b200 import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
res = len(sys.argv) > 5 and sys.argv[5] == "1"
bf = torch.bfloat16
x = (torch.randn(M, K, device="cuda") * 0.5).to(bf)
w = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
b = torch.zeros(N, device="cuda")
r = (torch.randn(M, N, device="cuda") * 0.5).to(bf) if res else None
out = torch.empty(M, N, device="cuda", dtype=bf)
for _ in range(4):
    ops.linear(x, w, b, residual=r, act=act, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
