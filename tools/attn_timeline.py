"""Timeline of one CTA of attn_tc2 (debug stamps): python tools/attn_timeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda", 0)
dbg = torch.zeros(3 * 64 * 8, device=dev, dtype=torch.int64)
os.environ["C2D_ATTN_DBG"] = hex(dbg.data_ptr())
from clap2diffusion_b200 import ops
B, N, C = 16, 4096, 320
qkv = (torch.randn(B, N, 3 * C, device=dev) * 0.5).to(torch.bfloat16)
out = torch.empty(B, N, C, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 8, out=out)
torch.cuda.synchronize()
d = dbg.cpu().view(3, 64, 8)
t0 = int(d[0, 0, 0])
print("softmax WG0 / WG1: tile: start s_full ld_done exp_start exp_end o_done p_full   (clk rel. to WG0 tile0 start)")
for j in list(range(0, 6)) + [30, 31]:
    for r in (0, 1):
        print(f"  WG{r} t{j:2d}: " + " ".join(f"{int(d[r, j, e]) - t0:7d}" for e in range(7)))
    print(f"  MMA t{j:2d}: s_free0 {int(d[2,j,0])-t0:7d} s_free1 {int(d[2,j,1])-t0:7d} qk_issued {int(d[2,j,2])-t0:7d} p_full0 {int(d[2,j,3])-t0:7d} p_full1 {int(d[2,j,4])-t0:7d} pv_issued {int(d[2,j,5])-t0:7d}")
per = (int(d[0, 31, 0]) - int(d[0, 1, 0])) / 30
print("mean tile period (WG0):", per, "clk")
