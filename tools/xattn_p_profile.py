#!/usr/bin/env python
"""Wait-cycle breakdown of CTA 0 of the persistent fused cross-attention kernel (xattn_tc2.cu; GPU only).

    python tools/xattn_p_profile.py [--batch 16]

Prints, per role, the cycles spent waiting on each mbarrier and in its compute sections (clock64 accumulators,
C2D_XATTN_DBG), next to the role's total -- the tool that says which pipeline edge is the critical path."""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from clap2diffusion_b200 import ops  # noqa: E402

ROLES = {1: ("projection issuer", ["x_full", "-", "-", "-", "-", "qbf_ready(n-1)", "w_full"]),
         2: ("attention issuer", ["-", "k_full", "p_full", "o_free", "v_full", "qbf_ready"]),
         4: ("softmax A (warp 4)", ["s_full", "compute"]), 8: ("softmax B (warp 8)", ["s_full", "compute"]),
         12: ("convert/epilogue (warp 12)", ["q_done", "qk_done", "o_full", "convert", "epilogue"])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    bf, B, heads, Nq, C = torch.bfloat16, a.batch, 8, 4096, 320
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, Nq, C, device=dev, generator=g).to(bf)
    kv = (torch.randn(B, 77, 2 * C, device=dev, generator=g) * 0.5).to(bf)
    w32 = torch.randn(C, C, device=dev, generator=g) * C ** -0.5
    ln = ops.pack_lnfold(w32, torch.ones(C, device=dev), torch.zeros(C, device=dev), None, bf)
    rs = torch.zeros(B * Nq * 2, device=dev, dtype=torch.int64)
    ops.linear(x, torch.eye(C, device=dev, dtype=bf), row_stats=rs)
    kvp = ops.xattn_pack_kv(kv, heads)
    out = torch.empty(B, Nq, C, device=dev, dtype=bf)
    for _ in range(3):
        ops.xattn(x, kvp, ln=ln, ln_stats=rs, out=out)
    torch.cuda.synchronize()
    dbg = torch.zeros(20 * 8, device=dev, dtype=torch.int64)
    os.environ["C2D_XATTN_DBG"] = hex(dbg.data_ptr())
    ops.xattn(x, kvp, ln=ln, ln_stats=rs, out=out)
    torch.cuda.synchronize()
    del os.environ["C2D_XATTN_DBG"]
    d = dbg.view(20, 8).cpu().tolist()
    items = (B * Nq // 128 * 2 + 147) // 148
    print(f"CTA 0: ~{items} items (128 rows x 4 heads each)")
    for wp, (name, ids) in ROLES.items():
        tot = d[wp][7]
        parts = ", ".join(f"{n} {d[wp][i]}" for i, n in enumerate(ids))
        print(f"  {name:28s} total {tot:7d} cycles ({tot / max(items, 1):7.0f} / item) | {parts}")




if __name__ == "__main__":
    main()
