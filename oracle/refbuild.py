"""TEST INFRASTRUCTURE (oracle/): build the UNMODIFIED reference (safaad/aim) natively.

The reference's sources are compiled where they lie under /root/reference with gcc and the
UPMEM stand-in headers in oracle/shim/ (the UPMEM SDK v2021.3.0 is a third-party dependency
that is neither vendored nor installed; it supplies transfers/launch only, no arithmetic).
Outputs go to oracle/_ref/ only (git-ignored, travels with gpurun).  No reference source is
copied into this repository.  Only tests/, __graft_entry__ and bench.py's CPU-baseline legs may
use this module; the product (aim_b200/) never does.

Each reference program is parametrised at COMPILE time (-DMAX_SCORE, -DREAD_SIZE, ... exactly
as */run-*-pim-*.py:133-139 passes them to make), so one binary exists per parameter set:
    oracle/_ref/<alg>-<mem>-<tag>            e.g. wfa-mram-ms30-rs168-x3-o4-e1-RB
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_ROOT = Path(os.environ.get("AIM_REFERENCE_ROOT", "/root/reference"))
OUT_DIR = HERE / "_ref"
SHIM = HERE / "shim"

_DIRS = {
    ("wfa", "mram"): "WFA/DPU-MRAM", ("wfa", "wram"): "WFA/DPU-WRAM",
    ("nw", "mram"): "NW/DPU-MRAM", ("nw", "wram"): "NW/DPU-WRAM",
    ("swg", "mram"): "SWG/DPU-MRAM", ("swg", "wram"): "SWG/DPU-WRAM",
    # aim-genasm submodule; the filter prints "idx, score", DC prints "idx, score, CIGAR"
    ("genasm_dc", "mram"): "aim-genasm/GenASM/DPU-MRAM-DC", ("genasm_dc", "wram"): "aim-genasm/GenASM/DPU-WRAM-DC",
    ("genasm_filter", "mram"): "aim-genasm/GenASM/DPU-MRAM-filter", ("genasm_filter", "wram"): "aim-genasm/GenASM/DPU-WRAM-filter",
}


def ref_name(alg: str, mem: str, *, max_score: int, read_size: int, match: int = 0, mismatch: int = 3,
             gap_o: int = 4, gap_e: int = 1, backtrace: bool = True, reduce: bool = False,
             nr_tasklets: int = 1, big_wram: bool = False) -> str:
    tag = f"{alg}-{mem}-ms{max_score}-rs{read_size}-m{match}-x{mismatch}-o{gap_o}-e{gap_e}-"
    tag += ("R" if reduce else "") + ("B" if backtrace else "") + ("W" if big_wram else "")
    if nr_tasklets != 1:
        tag += f"-t{nr_tasklets}"
    return tag


def ref_path(alg: str, mem: str, **kw) -> Path:
    return OUT_DIR / ref_name(alg, mem, **kw)


def reference_available() -> bool:
    return (REF_ROOT / "WFA" / "DPU-MRAM" / "dpu" / "wfa.c").exists()


def build_ref(alg: str, mem: str, *, max_score: int, read_size: int, match: int = 0, mismatch: int = 3,
              gap_o: int = 4, gap_e: int = 1, backtrace: bool = True, reduce: bool = False,
              nr_tasklets: int = 1, big_wram: bool = False, force: bool = False) -> Path:
    """Compile one reference program; returns the binary path (cached by name).

    For NW, gap_o is the single linear gap (GAP_I = GAP_D, NW/*/run-nw-pim-*.py:138-140).
    big_wram=True compiles a /tmp copy of dpu_allocator_wram.c whose 62000-byte guard
    (dpu_allocator_wram.c:6) is raised, needed only for WFA +BT at 10 kbp.
    """
    out = ref_path(alg, mem, max_score=max_score, read_size=read_size, match=match, mismatch=mismatch,
                   gap_o=gap_o, gap_e=gap_e, backtrace=backtrace, reduce=reduce,
                   nr_tasklets=nr_tasklets, big_wram=big_wram)
    if out.exists() and not force:
        return out
    if not reference_available():
        raise FileNotFoundError(f"reference tree not present at {REF_ROOT}; prebuilt {out.name} missing")
    src = REF_ROOT / _DIRS[(alg, mem)]
    OUT_DIR.mkdir(exist_ok=True)
    wram_segment = 60000 // nr_tasklets - 8
    defs = [f"-DMAX_SCORE={max_score}", f"-DREAD_SIZE={read_size}", f"-DMATCH={match}", f"-DMISMATCH={mismatch}",
            f"-DNR_TASKLETS={nr_tasklets}", "-DNR_DPUS=1"]
    if alg == "nw":
        defs += [f"-DGAP_I={gap_o}", f"-DGAP_D={gap_o}"]
    else:
        defs += [f"-DGAP_O={gap_o}", f"-DGAP_E={gap_e}"]
    if backtrace:
        defs.append("-DBACKTRACE")
    if reduce:
        defs.append("-DREDUCE")
    dpu_srcs = sorted(str(p) for p in (src / "dpu").glob("*.c"))
    with tempfile.TemporaryDirectory(prefix="aimref") as tmp:
        if big_wram:
            wram_segment = 16 << 20
            patched = Path(tmp) / "dpu_allocator_wram.c"
            text = (src / "dpu" / "dpu_allocator_wram.c").read_text().replace("62000", "2000000000")
            patched.write_text(text)
            dpu_srcs = [s for s in dpu_srcs if not s.endswith("dpu_allocator_wram.c")] + [str(patched)]
        defs.append(f"-DWRAM_SEGMENT={wram_segment}")
        common = ["gcc", "-O2", "-w", "-std=gnu11", f"-I{SHIM}", f"-I{src / 'common'}", f"-I{src / 'dpu'}"] + defs
        if alg.startswith("genasm"):
            common += ["-include", "limits.h"]  # genasmDC.c uses ULLONG_MAX; the UPMEM toolchain's headers pull limits.h in
        objs = []
        for i, s in enumerate(dpu_srcs):
            o = Path(tmp) / f"dpu{i}.o"
            subprocess.run(common + ["-Dmain=dpu_main", "-Dedit_cigar_print=dpu_edit_cigar_print", "-c", s, "-o", str(o)],
                           check=True)
            objs.append(str(o))
        ho = Path(tmp) / "host.o"
        subprocess.run(common + ["-c", str(src / "host" / "host.c"), "-o", str(ho)], check=True)
        so = Path(tmp) / "shim.o"
        subprocess.run(common + ["-c", str(SHIM / "shim.c"), "-o", str(so)], check=True)
        tmp_out = Path(tmp) / "bin"
        subprocess.run(["gcc", "-O2", "-o", str(tmp_out), str(ho), str(so)] + objs + ["-lpthread", "-lm"], check=True)
        shutil.copy2(tmp_out, out)
    return out


def run_ref(binary: Path, pairs_file: str | os.PathLike, out_file: str | os.PathLike, n: int, *,
            nr_dpus: int = 1, threads: int = 1, timeout: float | None = None) -> str:
    """Run `host <pairs> <out> <N>` of a reference build; returns its stdout (phase timers)."""
    env = dict(os.environ, AIM_SHIM_NR_DPUS=str(nr_dpus), AIM_SHIM_THREADS=str(threads))
    with tempfile.TemporaryDirectory(prefix="aimrun") as cwd:  # host.c drops a ./dpu-out side file
        p = subprocess.run([str(binary), str(Path(pairs_file).resolve()), str(Path(out_file).resolve()), str(n)],
                           cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference run failed rc={p.returncode}: {p.stdout[-400:]} {p.stderr[-400:]}")
    return p.stdout


def md5(path: str | os.PathLike) -> str:
    return hashlib.md5(Path(path).read_bytes()).hexdigest()


# Parameter sets of BASELINE.json's five configs + the CPU-baseline builds.
CONFIGS = {
    1: dict(alg="wfa", mem="mram", max_score=5, read_size=112, mismatch=3, gap_o=4, gap_e=1, backtrace=True),
    2: dict(alg="nw", mem="wram", max_score=4, read_size=112, mismatch=3, gap_o=4, backtrace=True),
    3: dict(alg="swg", mem="mram", max_score=80, read_size=272, match=0, mismatch=4, gap_o=6, gap_e=2, backtrace=True),
    4: dict(alg="wfa", mem="mram", max_score=30, read_size=168, mismatch=3, gap_o=4, gap_e=1, backtrace=True, reduce=True),
    5: dict(alg="wfa", mem="mram", max_score=5000, read_size=11008, mismatch=3, gap_o=4, gap_e=1, backtrace=False, reduce=True),
}


def build_all() -> list[Path]:
    outs = [build_ref(**cfg) for cfg in CONFIGS.values()]
    outs.append(build_ref("nw", "mram", max_score=4, read_size=112, mismatch=3, gap_o=4, backtrace=True))
    outs.append(build_ref("wfa", "mram", max_score=5000, read_size=11008, backtrace=True, reduce=True, big_wram=True))
    # aim-genasm: bench.py configs 7-9
    outs.append(build_ref("genasm_dc", "wram", max_score=5, read_size=112, backtrace=True))
    outs.append(build_ref("genasm_filter", "wram", max_score=1, read_size=112, backtrace=False))
    outs.append(build_ref("genasm_dc", "wram", max_score=30, read_size=168, backtrace=True))
    return outs


if __name__ == "__main__":
    for p in build_all():
        print(p)
