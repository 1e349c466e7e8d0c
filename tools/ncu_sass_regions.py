#!/usr/bin/env python
"""Per-SASS-instruction executed counts of an ncu capture (--page source --csv), printed as a listing with
cumulative shares, so that loop bodies can be read off by address range.

    python tools/ncu_sass_regions.py <report.ncu-rep> [start_hex end_hex]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = src.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
tot = sum(int(r["Instructions Executed"]) for r in rows)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
acc = 0
for r in rows:
    a = int(r["Address"], 16) - base
    n = int(r["Instructions Executed"])
    if lo <= a <= hi:
        acc += n
        print(f"{a:04x} {n/tot*100:5.2f}% thr {float(r['Avg. Threads Executed']):4.1f} smp {r['# Samples']:>6}  {r['Source'].strip()}")
print(f"range share {acc/tot*100:.2f}% of {tot} warp-inst")
