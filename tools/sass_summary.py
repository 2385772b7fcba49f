#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libc2d.so (no GPU needed): which kernels carry tcgen05 / TMA instructions.

    python tools/sass_summary.py > profiles/sass_summary_rN.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clap2diffusion_b200", "libc2d.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|FFMA2?|MUFU\.EX2)\b")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    rows = []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        rows.append((name, collections.Counter(m.group(1) for m in PAT.finditer(f))))
    names = subprocess.run(["c++filt"], input="\n".join(n for n, _ in rows), capture_output=True, text=True).stdout.split("\n")
    tot = collections.Counter()
    for _, c in rows:
        tot.update(c)
    print("# SASS summary of clap2diffusion_b200/libc2d.so (sm_100a); tools/sass_summary.py")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,")
    print("# UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit.  HMMA (legacy mma.sync) must be absent.")
    print(f"\nfunctions: {len(rows)}; totals: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items())) + f"; HMMA {tot.get('HMMA', 0)}\n")
    print("kernels that issue tcgen05.mma:")
    print(f"{'UTCHMMA':>8} {'.2CTA':>6} {'LDTM':>5} {'STTM':>5} {'UTMALDG':>8} {'UTMASTG':>8} {'UBLKCP':>7} {'UTCBAR':>7}  kernel")
    for (_, c), name in sorted(zip(rows, names), key=lambda r: r[1]):
        if c.get("UTCHMMA", 0) or c.get("UTCHMMA.2CTA", 0):
            short = re.sub(r"\(.*", "", name)
            print(f"{c.get('UTCHMMA', 0):8d} {c.get('UTCHMMA.2CTA', 0):6d} {c.get('LDTM', 0):5d} {c.get('STTM', 0):5d} {c.get('UTMALDG', 0):8d} "
                  f"{c.get('UTMASTG', 0):8d} {c.get('UBLKCP', 0):7d} {c.get('UTCBAR', 0):7d}  {short}")


if __name__ == "__main__":
    main()
