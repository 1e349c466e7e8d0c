#!/usr/bin/env python
"""List the SASS of a profiled kernel in program order with per-instruction executed counts and source lines.
    python tools/ncu_sass.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [min-share-of-max]"""
import csv, io, re, subprocess, sys
rep, cubin, kern = sys.argv[1:4]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(src) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(src[start:]))))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
infn, cur, insts = False, None, []
for l in dis:
    if re.match(r"\s*\.section\s+\.text\.", l) or l.startswith(".text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1][:14], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        insts.append(cur)
mx = max(int(r["Instructions Executed"]) for r in rows)
for loc, r in zip(insts, rows):
    ie = int(r["Instructions Executed"])
    if ie >= thr * mx:
        te = int(r["Thread Instructions Executed"])
        print(f"{ie/1e6:9.1f}M l{te/max(ie,1):4.0f} s{int(r['# Samples']):6d} {str(loc):24s} {r['Source'][:70]}")
