"""Run the one-pass GroupNorm a few times at the benchmark shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clap2diffusion_b200 import ops
dev = torch.device("cuda", 0)
B, HW, C = 16, 4096, 320
x = torch.randn(B, HW, C, device=dev).to(torch.bfloat16)
gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
s1 = ops.channel_stats(x, torch.zeros(B * C * 2, device=dev, dtype=torch.int64))
out = torch.empty_like(x)
for _ in range(5):
    ops.group_norm_apply(x, s1, gam, bet, 32, 1e-5, True, out=out)
torch.cuda.synchronize()
print("ok")
