#!/usr/bin/env python
"""Opcode mix (alu-pipe vs fma-pipe instructions) of an address range of a kernel's SASS.
    python tools/scan_mix.py <object> <mangled-substring> <start-hex> <end-hex>"""
import collections
import subprocess
import sys

obj, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["bash", "tools/sass_fn.sh", obj, pat], capture_output=True, text=True).stdout
c = collections.Counter()
FMA = {"IMAD", "IMAD.IADD", "IMAD.MOV", "IMAD.HI", "IMAD.U32", "IMAD.SHL", "IMAD.WIDE", "IMAD.X", "IDP"}
for line in txt.splitlines():
    f = line.split()
    a = int(f[0], 16)
    if not lo <= a <= hi:
        continue
    op = f[2] if f[1].startswith("@") else f[1]
    p = op.rstrip(";").split(".")
    name = p[0] + ("." + p[1] if p[0] == "IMAD" and len(p) > 1 and p[1] in ("IADD", "MOV", "HI", "U32", "SHL", "WIDE", "X") else "")
    c[name] += 1
tot = sum(c.values())
fma = sum(v for k, v in c.items() if k in FMA)
other = sum(v for k, v in c.items() if k in ("SHFL", "LDG", "STG", "BRA", "BSSY", "BSYNC", "LDC", "WARPSYNC", "VOTE", "NOP", "S2R", "LDS", "STS", "CREDUX", "R2UR"))
print(" ".join(f"{k}:{v}" for k, v in c.most_common()))
print(f"total {tot}  fma-pipe {fma}  alu-pipe ~{tot - fma - other}  other {other}")
