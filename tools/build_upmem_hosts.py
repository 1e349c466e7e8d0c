#!/usr/bin/env python
"""Compile one of safaad/aim's six host programs UNCHANGED against include/dpu.h + libaim_dpu.so (INTEGRATION.md route C).

    python tools/build_upmem_hosts.py --alg wfa --mem mram --max-score 30 --read-size 168 -b -r [-x 3 -g 4 -a 1 -m 0]
                                      [--nr-dpus 1] [--reference /root/reference] [-o build/upmem_hosts/<name>]

This is what `make` in a reference program directory does for its host (<prog>/Makefile:26-27,44: $(CC) host/host.c
`dpu-pkg-config --cflags --libs dpu` -DNR_DPUS -DNR_TASKLETS $(FLAGS)), with the UPMEM SDK's include/link flags replaced by
-Iinclude -laim_dpu.  The DPU binary is not built: the B200 library is the DPU program.  The reference source is read where
it lies; only the compiled host lands under build/ (git-ignored).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DIRS = {("wfa", "mram"): "WFA/DPU-MRAM", ("wfa", "wram"): "WFA/DPU-WRAM", ("nw", "mram"): "NW/DPU-MRAM",
        ("nw", "wram"): "NW/DPU-WRAM", ("swg", "mram"): "SWG/DPU-MRAM", ("swg", "wram"): "SWG/DPU-WRAM"}


def host_name(alg, mem, *, max_score, read_size, match=0, mismatch=3, gap_o=4, gap_e=1, backtrace=True, reduce=False, nr_dpus=1):
    tag = f"{alg}-{mem}-ms{max_score}-rs{read_size}-m{match}-x{mismatch}-o{gap_o}-e{gap_e}-" + ("R" if reduce else "") + ("B" if backtrace else "")
    return tag + (f"-d{nr_dpus}" if nr_dpus != 1 else "")


def build_host(alg, mem, *, max_score, read_size, match=0, mismatch=3, gap_o=4, gap_e=1, backtrace=True, reduce=False,
               nr_dpus=1, nr_tasklets=1, reference=None, out=None, force=False) -> Path:
    """-> path of the compiled host (`<out> <pairs> <out-file> <N>`, exactly the reference's command line)."""
    reference = Path(reference or os.environ.get("AIM_REFERENCE_ROOT", "/root/reference"))
    out = Path(out) if out else ROOT / "build" / "upmem_hosts" / host_name(
        alg, mem, max_score=max_score, read_size=read_size, match=match, mismatch=mismatch, gap_o=gap_o, gap_e=gap_e,
        backtrace=backtrace, reduce=reduce, nr_dpus=nr_dpus)
    lib = ROOT / "aim_b200" / "libaim_dpu.so"
    if out.exists() and not force and out.stat().st_mtime >= max(lib.stat().st_mtime, (ROOT / "include" / "dpu.h").stat().st_mtime):
        return out
    src = reference / DIRS[(alg, mem)]
    host_c = src / "host" / "host.c"
    if not host_c.exists():
        raise FileNotFoundError(f"{host_c}: reference tree not present")
    defs = [f"-DNR_DPUS={nr_dpus}", f"-DNR_TASKLETS={nr_tasklets}", f"-DMAX_SCORE={max_score}", f"-DREAD_SIZE={read_size}",
            f"-DMATCH={match}", f"-DMISMATCH={mismatch}"]
    defs += [f"-DGAP_I={gap_o}", f"-DGAP_D={gap_o}"] if alg == "nw" else [f"-DGAP_O={gap_o}", f"-DGAP_E={gap_e}"]
    if backtrace:
        defs.append("-DBACKTRACE")
    if reduce:
        defs.append("-DREDUCE")
    out.parent.mkdir(parents=True, exist_ok=True)
    libdir = ROOT / "aim_b200"
    cmd = ["gcc", "-O2", "-w", "-std=gnu11", f"-I{ROOT / 'include'}", f"-I{src / 'common'}", *defs, str(host_c), "-o", str(out),
           f"-L{libdir}", "-laim_dpu", "-laim_b200", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,$ORIGIN/../../aim_b200", "-lm"]
    subprocess.run(cmd, check=True)
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--alg", required=True, choices=["wfa", "nw", "swg"])
    ap.add_argument("--mem", default="mram", choices=["mram", "wram"])
    ap.add_argument("--max-score", type=int, required=True)
    ap.add_argument("--read-size", type=int, required=True)
    ap.add_argument("-m", "--match", type=int, default=0)
    ap.add_argument("-x", "--mismatch", type=int, default=3)
    ap.add_argument("-g", "--gap-o", type=int, default=4)
    ap.add_argument("-a", "--gap-e", type=int, default=1)
    ap.add_argument("-b", "--backtrace", action="store_true")
    ap.add_argument("-r", "--reduce", action="store_true")
    ap.add_argument("--nr-dpus", type=int, default=1)
    ap.add_argument("--reference", default=None)
    ap.add_argument("-o", "--out", default=None)
    a = ap.parse_args()
    print(build_host(a.alg, a.mem, max_score=a.max_score, read_size=a.read_size, match=a.match, mismatch=a.mismatch, gap_o=a.gap_o,
                     gap_e=a.gap_e, backtrace=a.backtrace, reduce=a.reduce, nr_dpus=a.nr_dpus, reference=a.reference, out=a.out, force=True))


if __name__ == "__main__":
    sys.exit(main())
