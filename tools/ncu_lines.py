#!/usr/bin/env python
"""Attribute an ncu capture's per-SASS-instruction counters to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]

ncu's `--page source --csv` gives per-SASS-instruction counts; `nvdisasm -g` gives the source line of
every SASS instruction of the same cubin.  Joined by instruction order inside the kernel.
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    rep, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = src.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    if cubin.endswith(".o"):  # a host object with an embedded cubin: extract it first (nvdisasm reads ELF cubins only)
        import os, tempfile
        tmp = tempfile.mkdtemp(prefix="ncu_lines")
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(cubin)], cwd=tmp, capture_output=True)
        found = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        if found:
            cubin = os.path.join(tmp, found[0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # find the kernel's function body
    infn = False
    cur = None
    insts = []
    for l in dis:
        if l.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", l):
            infn = kern in l
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            insts.append((cur, l.split("*/", 1)[1].strip().rstrip(";")))
    if len(insts) != len(rows):
        sys.exit(f"error: {len(insts)} disassembled vs {len(rows)} profiled instructions - the kernel substring must select exactly "
                 "the profiled template instantiation of the same build")
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for (loc, _), r in zip(insts, rows):
        ie, te, sm = int(r["Instructions Executed"]), int(r["Thread Instructions Executed"]), int(r["# Samples"])
        a = agg[loc]
        a[0] += ie; a[1] += te; a[2] += sm
        tot[0] += ie; tot[1] += te; tot[2] += sm
    print(f"total warp-inst {tot[0]:,}  thread-inst {tot[1]:,}  samples {tot[2]:,}")
    srcs = {}
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        txt = ""
        if loc:
            try:
                if loc[0] not in srcs:
                    srcs[loc[0]] = open(f"/root/repo/aim_b200/csrc/{loc[0]}").read().splitlines()
                txt = srcs[loc[0]][loc[1] - 1].strip()[:90]
            except Exception:
                pass
        print(f"{100*a[0]/tot[0]:5.1f}% inst {100*a[2]/max(tot[2],1):5.1f}% samp  lanes {a[1]/max(a[0],1):4.1f}  {loc}  {txt}")


if __name__ == "__main__":
    main()
