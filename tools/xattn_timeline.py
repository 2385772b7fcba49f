"""Timeline of one mid-grid CTA of xattn_tc_kernel (debug stamps): python tools/xattn_timeline.py [site 0..3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda", 0)
dbg = torch.zeros(3 * 64, device=dev, dtype=torch.int64)
os.environ["C2D_XATTN_DBG"] = hex(dbg.data_ptr())
from clap2diffusion_b200 import ops
site = int(sys.argv[1]) if len(sys.argv) > 1 else 0
hw, C = [(64, 320), (32, 640), (16, 1280), (8, 1280)][site]
B, Nq, heads = 16, hw * hw, 8
bf = torch.bfloat16
x = torch.randn(B, Nq, C, device=dev).to(bf)
kv = (torch.randn(B, 77, 2 * C, device=dev) * 0.5).to(bf)
w = (torch.randn(C, C, device=dev) * C ** -0.5).to(bf)
out = torch.empty(B, Nq, C, device=dev, dtype=bf)
kvp = ops.xattn_pack_kv(kv, heads)
for _ in range(3):
    ops.xattn(x, kvp, wq=w, out=out)
torch.cuda.synchronize()
d = dbg.cpu().view(3, 64)
t0 = int(d[0, 0])
names = ["producer (start, ring slot free x num_kb, q_done, kv slot free x G)",
         "mma (ring full x num_kb, qbf_ready, then per head: kv_full, p_full+o_free)",
         "compute (setup done, q_done, conv iters.., qbf written, then per head: s_full, ld done, max done, exp done, p written, o_full)"]
for r in range(3):
    print(names[r])
    print("   ", " ".join(f"{int(v) - t0:6d}" for v in d[r] if int(v) != 0))
