"""Run the 4096-token self-attention a few times at the benchmark shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from clap2diffusion_b200 import ops
dev = torch.device("cuda", 0)
B, N, C = 16, int(os.environ.get("ATTN_N", "4096")), int(os.environ.get("ATTN_C", "320"))
qkv = (torch.randn(B, N, 3 * C, device=dev) * 0.5).to(torch.bfloat16)
out = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
for _ in range(4):
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 8, out=out)
torch.cuda.synchronize()
print("ok")
