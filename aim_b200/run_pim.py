"""Shared body of the six run-*-pim-*.py wrappers (scripts/), option-compatible with the reference's
{WFA,SWG,NW}/DPU-{WRAM,MRAM}/run-*-pim-*.py: same flags (-i -o -l -e -n -m -x -g -a -b -r -t -d), same
validation messages and exit codes, same MAX_SCORE/READ_SIZE derivation (run-wfa-pim-mram.py:58-67).
Where the reference runs `make ... FLAGS="-D..."` and then `./build/host in out N`, the knobs are
exported as environment variables of the same names and the B200 `build/host in out N` is exec'ed.
NR_TASKLETS/WRAM_SEGMENT (the DPU sizing heuristic, :70-118) have no meaning on a GPU: -t is accepted
and ignored, and the occupancy-driven launch configuration is chosen inside the library."""
from __future__ import annotations

import argparse
import math
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def build_parser(algo: str) -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(add_help=True)
    ap.add_argument("-i", "--input", type=str, required=True, help="Input read pairs file path")
    ap.add_argument("-o", "--output", type=str, help="Output alignment file path", default="./out")
    ap.add_argument("-l", "--read_length", required=True, type=int, help="Read length")
    ap.add_argument("-e", "--error", type=float, required=True, help="Percentage error per read length")
    ap.add_argument("-n", "--number_reads", type=int, required=True, help="Number of read pairs to be aligned")
    ap.add_argument("-m", "--match_cost", type=int, default=0, help="Cost of characters match")
    ap.add_argument("-x", "--mismatch_cost", type=int, default=3, help="Cost of characters mismatch")
    if algo == "nw":
        ap.add_argument("-g", "--gap", type=int, default=4, help="Cost of a new gap deletion/insertion")
    else:
        ap.add_argument("-g", "--gap_opening", type=int, default=4, help="Cost of opening a new gap")
        ap.add_argument("-a", "--gap_extending", type=int, default=1, help="Cost of extending gap")
    ap.add_argument("-b", "--backtrace", action="store_true", help="Enable backtracing")
    if algo == "wfa":
        ap.add_argument("-r", "--reduced", action="store_true", help="Enable WFA-Adaptive")
    ap.add_argument("-t", "--nr_of_tasklets", type=int, help="NR_TASKLETS (accepted, ignored on GPU)")
    ap.add_argument("-d", "--nr_of_dpus", type=int, help="NR_DPUs (only feeds the pairs-to-process rule; default=1)")
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("AIM_NGPUS", "1")), help="B200s to shard over (extension)")
    ap.add_argument("--dry-run", action="store_true", help="print the environment and command, do not run (extension)")
    return ap


def derive(algo: str, args: dict) -> dict:
    """The -D knob set the reference would compile with, as environment variables."""
    match_cost, mismatch_cost = args["match_cost"], args["mismatch_cost"]
    if algo == "nw":
        gap = args["gap"]
        if match_cost > 0 or mismatch_cost <= 0 or gap <= 0:
            print("Wrong affine gap penalties must be  m <= 0 and g, a, x > 0\n")
            sys.exit(-1)
    else:
        gap_opening, gap_extending = args["gap_opening"], args["gap_extending"]
        if match_cost > 0 or mismatch_cost <= 0 or gap_opening <= 0 or gap_extending <= 0:
            print("Wrong affine gap penalties must be  m <= 0 and g, a, x > 0\n")
            sys.exit(-1)
    read_length = args["read_length"]
    if read_length <= 0:
        print("Undefined input read length")
        sys.exit(-1)
    if args["number_reads"] <= 0:
        print("Undefined number of input reads")
        sys.exit(-1)
    nr_of_wrong_bases = read_length * args["error"]
    if algo == "nw":
        max_score = math.ceil(max(nr_of_wrong_bases * mismatch_cost, nr_of_wrong_bases * gap))
    else:
        max_score = math.ceil(max(nr_of_wrong_bases * mismatch_cost, nr_of_wrong_bases * (gap_opening + gap_extending)))
    read_size = math.ceil((((read_length + nr_of_wrong_bases) + 7) / 8)) * 8
    env = {"MAX_SCORE": int(max_score), "READ_SIZE": int(read_size), "MATCH": match_cost, "MISMATCH": mismatch_cost,
           "BACKTRACE": int(bool(args["backtrace"])), "REDUCE": int(bool(args.get("reduced", False))),
           "NR_DPUS": args["nr_of_dpus"] or 1, "AIM_ALGO": algo, "AIM_NGPUS": args["gpus"]}
    if algo == "nw":
        env.update(GAP_I=gap, GAP_D=gap)
    else:
        env.update(GAP_O=gap_opening, GAP_E=gap_extending)
    if args["nr_of_tasklets"] is not None:
        env["NR_TASKLETS"] = args["nr_of_tasklets"]
    return env


def main(algo: str, variant: str, argv=None) -> int:
    args = vars(build_parser(algo).parse_args(argv))
    env = derive(algo, args)
    env["AIM_VARIANT"] = variant
    host = Path(os.environ.get("AIM_HOST_BINARY", ROOT / "build" / "host"))
    flags = " ".join(f"-D{k}={v}" for k, v in env.items() if not k.startswith("AIM_") and k not in ("BACKTRACE", "REDUCE"))
    flags += (" -DREDUCE" if env["REDUCE"] else "") + (" -DBACKTRACE" if env["BACKTRACE"] else "")
    print(f"B200 runtime knobs (reference: make FLAGS): {flags}")
    cmd = [str(host), args["input"], args["output"], str(args["number_reads"])]
    print(" ".join(cmd))
    if args["dry_run"]:
        for k, v in sorted(env.items()):
            print(f"{k}={v}")
        return 0
    if not host.exists():
        print(f"{host} not found: run `make` first", file=sys.stderr)
        return 1
    return subprocess.call(cmd, env=dict(os.environ, **{k: str(v) for k, v in env.items()}))


# ---- aim-genasm wrappers (aim-genasm/GenASM/DPU-{WRAM,MRAM}-{DC,filter}/run-genasm{dc,filter}-pim-{wram,mram}.py) ----
def build_parser_genasm() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(add_help=True)
    ap.add_argument("-i", "--input", type=str, required=True, help="Input read pairs file path")
    ap.add_argument("-o", "--output", type=str, help="Output alignment file path", default="./out")
    ap.add_argument("-l", "--read_length", required=True, type=int, help="Read Length")
    ap.add_argument("-e", "--error", type=float, help="Percentage error per read length (or provide max edit distance)")
    ap.add_argument("-n", "--number_reads", type=int, required=True, help="Number of read pairs to be aligned")
    ap.add_argument("-m", "--match_cost", type=int, default=0, help="Cost of characters match")
    ap.add_argument("-x", "--mismatch_cost", type=int, default=3, help="Cost of characters mismatch")
    ap.add_argument("-g", "--gap_opening", type=int, default=4, help="Cost of opening a new gap")
    ap.add_argument("-a", "--gap_extending", type=int, default=1, help="Cost of extending gap")
    ap.add_argument("-k", "--max_edit", type=int, help="max edit distance operations (optional or provide percentage error)")
    ap.add_argument("-t", "--nr_of_tasklets", type=int, help="NR_TASKLETS (accepted, ignored on GPU)")
    ap.add_argument("-d", "--nr_of_dpus", type=int, help="NR_DPUs (only feeds the pairs-to-process rule; default=1)")
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("AIM_NGPUS", "1")), help="B200s to shard over (extension)")
    ap.add_argument("--dry-run", action="store_true", help="print the environment and command, do not run (extension)")
    return ap


def derive_genasm(kind: str, args: dict) -> dict:
    """MAX_SCORE / READ_SIZE as run-genasmdc-pim-wram.py:36-70 (DC) and run-genasmfilter-pim-wram.py:36-67 (filter) derive them."""
    match_cost, mismatch_cost = args["match_cost"], args["mismatch_cost"]
    gap_opening, gap_extending = args["gap_opening"], args["gap_extending"]
    if match_cost > 0 or mismatch_cost <= 0 or gap_opening <= 0 or gap_extending <= 0:
        print("Wrong affine gap penalties must be  m <= 0 and g, a, x > 0\n")
        sys.exit(-1)
    read_length = args["read_length"]
    if read_length <= 0:
        print("Undefined input read length")
        sys.exit(-1)
    if args["number_reads"] <= 0:
        print("Undefined number of input reads")
        sys.exit(-1)
    if args["max_edit"] is not None:
        max_score = nr_of_wrong_bases = args["max_edit"]
    elif args["error"] is not None:
        if kind == "dc":
            nr_of_wrong_bases = read_length * args["error"]
            max_score = math.ceil(max(nr_of_wrong_bases * mismatch_cost, nr_of_wrong_bases * (gap_opening + gap_extending)))
        else:
            nr_of_wrong_bases = math.ceil(read_length * args["error"])
            max_score = nr_of_wrong_bases
    else:
        print("Missing input provide either max_edit or \\%\\ error" if kind == "dc"
              else "missing input provide either max number of edits or \\%\\ error")
        sys.exit(-1)
    if nr_of_wrong_bases == 0:
        nr_of_wrong_bases = 1
        max_score = 1
    read_size = math.ceil((((read_length + nr_of_wrong_bases) + 7) / 8)) * 8
    env = {"MAX_SCORE": int(max_score), "READ_SIZE": int(read_size), "MATCH": match_cost, "MISMATCH": mismatch_cost,
           "GAP_O": gap_opening, "GAP_E": gap_extending, "NR_DPUS": args["nr_of_dpus"] or 1,
           "AIM_ALGO": "genasm_dc" if kind == "dc" else "genasm_filter", "AIM_NGPUS": args["gpus"]}
    if args["nr_of_tasklets"] is not None:
        env["NR_TASKLETS"] = args["nr_of_tasklets"]
    return env


def main_genasm(kind: str, variant: str, argv=None) -> int:
    args = vars(build_parser_genasm().parse_args(argv))
    env = derive_genasm(kind, args)
    env["AIM_VARIANT"] = variant
    host = Path(os.environ.get("AIM_HOST_BINARY", ROOT / "build" / "host"))
    flags = " ".join(f"-D{k}={v}" for k, v in env.items() if not k.startswith("AIM_"))
    print(f"B200 runtime knobs (reference: make FLAGS): {flags}")
    cmd = [str(host), args["input"], args["output"], str(args["number_reads"])]
    print(" ".join(cmd))
    if args["dry_run"]:
        for k, v in sorted(env.items()):
            print(f"{k}={v}")
        return 0
    if not host.exists():
        print(f"{host} not found: run `make` first", file=sys.stderr)
        return 1
    return subprocess.call(cmd, env=dict(os.environ, **{k: str(v) for k, v in env.items()}))
