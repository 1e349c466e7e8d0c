#!/usr/bin/env python
"""bench.py — aligned pairs/s (score + CIGAR) of the B200 aligner on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 4] [--pairs P]

Workload (config.workload): BASELINE.json configs[3] = "WFA-adaptive l=150 e=4 %, 10M synthetic pairs
with backtrace" (the configuration the north-star target is quoted on; it fits one GPU).  One STEP =
one pass of the hot path over the whole batch of P pairs held by a rank.  Pairs shard by index with
no collective, so N ranks each align their own P pairs ("scaling": "weak") and
value = N*P / max-over-ranks(step time).

  value     device-resident: inputs already in HBM, kernels only, CUDA events on the launch stream.
  e2e       the same batch through the reference-facing C-ABI call aim_align_batch() with PINNED
            HOST buffers: H2D of sequences, alignment, D2H of results + CIGAR ops inside the timed
            region (double-buffered streams inside the library).
  roofline  HBM view of the dominant kernel (schema of the task) + int_roofline: the kernel is
            INT32-ALU bound (integer wavefront DP, no tensor cores), so the meaningful ceiling is the
            measured integer rate; both are reported, see DESIGN.md.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref, native build of the DPU C sources) on the
            host cores, on a bounded sample of the same workload.

--impl reference times that reference build alone, with all host threads, on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# BASELINE.json configs: index -> workload description + knobs
CONFIGS = {
    2: dict(name="NW linear-gap l=100 e=1% score+CIGAR", algo="nw", length=100, error=0.01, mismatch=3, gap_open=4, gap_ext=1,
            reduce=False, backtrace=True, pairs=2_000_000, seed=2),
    3: dict(name="SWG gap-affine x4 g6 a2 l=250 e=4% 1M pairs +BT", algo="swg", length=250, error=0.04, mismatch=4, gap_open=6,
            gap_ext=2, reduce=False, backtrace=True, pairs=1_000_000, seed=3),
    4: dict(name="WFA-adaptive l=150 e=4% 10M synthetic pairs +BT", algo="wfa", length=150, error=0.04, mismatch=3, gap_open=4,
            gap_ext=1, reduce=True, backtrace=True, pairs=10_000_000, seed=4),
    5: dict(name="WFA-adaptive long reads l=10000 e=10% 200K pairs score-only", algo="wfa", length=10000, error=0.10, mismatch=3,
            gap_open=4, gap_ext=1, reduce=True, backtrace=False, pairs=200_000, seed=5),
    # config 5 WITH backtrace: beyond what the reference can run, informational
    6: dict(name="WFA-adaptive long reads l=10000 e=10% +BT (beyond the reference's limits)", algo="wfa", length=10000, error=0.10,
            mismatch=3, gap_open=4, gap_ext=1, reduce=True, backtrace=True, pairs=50_000, seed=5),
    # SURVEY 8f item 3 (aim-genasm submodule): informational, not a BASELINE.json config
    7: dict(name="GenASM-DC l=100 e=1% (k=5 error levels) synthetic pairs, score + CIGAR string", algo="genasm_dc", length=100, error=0.01,
            mismatch=3, gap_open=4, gap_ext=1, reduce=False, backtrace=True, pairs=10_000_000, seed=7),
    8: dict(name="GenASM-filter l=100 e=1% (k=1 edit) synthetic pairs, edit-distance filter", algo="genasm_filter", length=100, error=0.01,
            mismatch=3, gap_open=4, gap_ext=1, reduce=False, backtrace=False, pairs=10_000_000, seed=7),
    9: dict(name="GenASM-DC l=150 e=4% (k=30 error levels) synthetic pairs, score + CIGAR string", algo="genasm_dc", length=150, error=0.04,
            mismatch=3, gap_open=4, gap_ext=1, reduce=False, backtrace=True, pairs=2_000_000, seed=4),
}
# dominant kernel of each config (the one `roofline` describes; one step = all kernels of the launch)
KERNELS = {2: "dp2_strip_kernel<NW> (two pairs per thread) + dp_row_kernel<NW> (aim_dp_fast.cu, aim_dp_pack2.cuh)", 3: "dp2_strip_kernel<SWG> (two pairs per thread) + dp_scan_kernel<SWG, 16 columns x 2 per lane, 8 lanes per pair, traceback by the warp's lanes> for the aliased pairs (aim_dp_fast.cu, aim_dp_pack2.cuh, aim_dp_scan.cuh)",
           4: "wfa_sub_kernel<4, reduce, backtrace, narrow rows> (aim_wfa_sub.cu)", 5: "wfa_long_kernel<16, reduce, 128- then 256-diagonal window> (aim_wfa_long.cu)",
           6: "wfa_long_kernel<16, reduce, backtrace, 128- then 256-diagonal window> (aim_wfa_long.cu)", 7: "genasm_band_kernel<1 word, DC> + genasm_tb_kernel (aim_genasm.cu)",
           8: "genasm_band_kernel<1 word, filter> (aim_genasm.cu)", 9: "genasm_band_kernel<4 words, DC> + genasm_tb_kernel (aim_genasm.cu)"}
INT_OPS_PER_GENASM_WORD = 14  # one (text step, level, 64-bit word): 3 shifts with carry, 1 or, 3 and = 7 64-bit ops = 14 int32 ops


def knobs(cfg):
    """(MAX_SCORE, READ_SIZE) of a config as the reference's run scripts derive them, in the same Python float
    arithmetic (run-wfa-pim-mram.py:58-67, run-nw-pim-mram.py:51-60).  Pure Python: the reference arm must not load
    product code; the GPU arm asserts that aim_derive_knobs (C ABI) agrees."""
    import math
    l, e = cfg["length"], cfg["error"]
    if cfg["algo"] == "genasm_filter":  # run-genasmfilter-pim-wram.py:57-67
        w = math.ceil(l * e) or 1
        return w, math.ceil((l + w + 7) / 8) * 8
    w = l * e
    gap = w * cfg["gap_open"] if cfg["algo"] == "nw" else w * (cfg["gap_open"] + cfg["gap_ext"])
    # run-genasmdc-pim-wram.py:57-70 = the WFA script's formula
    return math.ceil(max(w * cfg["mismatch"], gap)), math.ceil((l + w + 7) / 8) * 8


def config_dict(cfg, P, world):
    """`config` of the JSON line: identical for the GPU arm and the reference arm (the driver compares them)."""
    ms, rs = knobs(cfg)
    return {"workload": cfg["name"], "pairs_per_gpu_per_step": P, "algo": cfg["algo"], "max_score": ms, "read_size": rs,
            "penalties": {"x": cfg["mismatch"], "o": cfg["gap_open"], "e": cfg["gap_ext"]}, "backtrace": cfg["backtrace"],
            "adaptive": cfg["reduce"], "parallelism": f"pairs sharded by index over {world} GPU(s), no collective",
            "l2": f"inputs {P * 2 * rs / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
            "generator": f"seed {cfg['seed']}, generate_dataset semantics"}
# bounded CPU-reference sample per config (about 10-30 s of CPU work on 16 threads)
REF_SAMPLE = {4: 2_000_000, 2: 400_000, 3: 16_000, 5: 160, 6: 160, 7: 1_000_000, 8: 2_000_000, 9: 100_000}
# pairs per rank checked bit-exact against the oracle after the timed region (0 = every pair of the step); long reads and
# GenASM take a bounded, evenly spread sample (the oracle does ~1e2 long-read pairs/s per core)
PARITY_SAMPLE = {2: 0, 3: 0, 4: 0, 5: 20_000, 6: 2_000, 7: 1_000_000, 8: 2_000_000, 9: 100_000}
# algorithmic work per pair (restated in DESIGN.md "Measurement")
INT_OPS_PER_OFFSET = 11   # one computed (score, diagonal) offset: I, D, M recurrences
INT_OPS_PER_EXTEND = 4    # xor, clz, add, cmp per 16-base word step


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_run(cfg: dict, pairs: int, threads: int, repeats: int = 1, warmup: int = 0) -> dict:
    """Time the UNMODIFIED reference (native build of its DPU + host C sources, oracle/_ref) on the host
    cores: one simulated DPU per chunk of pairs, `threads` host threads.  The pair file is written once;
    the reference host is run warmup + repeats times and the phase timers it prints are averaged."""
    from oracle import refbuild as rb
    ms, rs = knobs(cfg)
    kw = dict(max_score=ms, read_size=rs, mismatch=cfg["mismatch"], gap_o=cfg["gap_open"], gap_e=cfg["gap_ext"],
              backtrace=cfg["backtrace"], reduce=cfg["reduce"])
    mem = "wram" if cfg["algo"] in ("nw", "genasm_dc", "genasm_filter") else "mram"
    binary = rb.build_ref(cfg["algo"], mem, **kw)  # prebuilt in oracle/_ref on the GPU box
    # one DPU image holds 64 MB: keep each simulated DPU well below it (host.c:215-241 layout)
    per_pair = 8 + 32 + 2 * rs + (2 * rs if cfg["backtrace"] else 0)
    hist = (rs * rs * 8) if cfg["algo"] == "swg" else (rs * rs * 2 if cfg["algo"] == "nw" else 4 << 20)
    if cfg["algo"] == "genasm_dc":  # traceback matrix per tasklet: n * (k+1) * 4 * count words (genasmDC.c:386)
        hist = rs * (ms + 1) * 4 * ((rs + 64) // 64) * 8 + (1 << 20)
    max_per_dpu = max(8, ((60_000_000 - hist) // per_pair) // 8 * 8)
    nr_dpus = max(threads, -(-pairs // max_per_dpu))
    nr_dpus = -(-nr_dpus // threads) * threads
    ph_sum, runs, lines, wall = [0.0, 0.0, 0.0], 0, 0, 0.0
    with tempfile.TemporaryDirectory(prefix="aimbench") as tmp:
        pairs_file = Path(tmp) / "in.pairs"
        # the stand-alone generator (tools/genpairs.cpp, host-only code): no product library in this process tree
        subprocess.run([str(ROOT / "build" / "aim_genpairs"), str(cfg["seed"]), "0", str(pairs), str(cfg["length"]), repr(cfg["error"]),
                        str(rs), str(pairs_file)], check=True)
        for it in range(warmup + repeats):
            t0 = time.perf_counter()
            out = rb.run_ref(binary, pairs_file, Path(tmp) / "out", pairs + 8 * nr_dpus, nr_dpus=nr_dpus, threads=threads)
            if it < warmup:
                continue
            wall += time.perf_counter() - t0
            ph = [float(x) for x in re.findall(r"(?:CPU-DPU|DPU Kernel|DPU-CPU)(?: Time)?: ([0-9.]+) ms", out)]
            ph_sum = [a + b for a, b in zip(ph_sum, ph)]
            runs += 1
        lines = (Path(tmp) / "out").read_bytes().count(b"\n")
    aligned = lines // (2 if cfg["backtrace"] else 1)
    return dict(pairs=aligned, h2d_ms=ph_sum[0] / runs, kernel_ms=ph_sum[1] / runs, d2h_ms=ph_sum[2] / runs, wall_s=wall / runs,
                nr_dpus=nr_dpus, threads=threads, binary=binary.name)


def run_reference_arm(args, cfg) -> None:
    rank, _, world = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = args.ref_pairs or REF_SAMPLE[args.config]
    last = cpu_reference_run(cfg, sample, threads, repeats=args.steps, warmup=args.warmup)
    t = (last["h2d_ms"] + last["kernel_ms"] + last["d2h_ms"]) * 1e-3
    value = last["pairs"] / t
    ms, rs = knobs(cfg)
    line = {
        "impl": "reference", "metric": "aligned pairs/sec (score+CIGAR)", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64" if cfg["algo"].startswith("genasm") else "int16", "data": "synthetic",
        "config": config_dict(cfg, args.pairs or cfg["pairs"], max(1, world)),
        "sample_pairs_per_step": last["pairs"],
        "note": "reference DPU C sources compiled natively (UPMEM SDK stand-in), one host thread per simulated DPU; "
                "time = its own CPU-DPU + DPU Kernel + DPU-CPU phases; UPMEM functional simulator unavailable (not installed)",
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference",
                         "sample": f"{last['pairs']} pairs of the workload per step, {last['nr_dpus']} simulated DPUs on {threads} threads",
                         "kernel_only_value": last["pairs"] / (last["kernel_ms"] * 1e-3)},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per rank per step (default: the config's)")
    ap.add_argument("--ref-pairs", type=int, default=0, help="CPU reference sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--parity", default="auto", choices=["auto", "full", "off"],
                    help="bit-exact check of the step's results against the CPU oracle, outside the timed region "
                         "(auto: every pair for configs 2-4, a bounded sample for long reads / GenASM)")
    ap.add_argument("--parity-pairs", type=int, default=-1, help="pairs per rank to check (0 = all)")
    ap.add_argument("--no-cli", action="store_true", help="skip the pair-file -> output-file leg (aim_align_file)")
    ap.add_argument("--no-inproc", action="store_true", help="skip the one-process aim_align_batch(ngpus=N) leg at N>1")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    rank, local_rank, world = dist_env()
    # stdout carries exactly ONE JSON line: library chatter (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist

    import aim_b200 as A
    from aim_b200 import shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; aim_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    P = args.pairs or cfg["pairs"]
    ms, rs = knobs(cfg)
    if not cfg["algo"].startswith("genasm"):
        assert (ms, rs) == A.derive_knobs(cfg["algo"], cfg["length"], cfg["error"], cfg["mismatch"], cfg["gap_open"], cfg["gap_ext"])
    params = A.AlignParams(algo=cfg["algo"], mismatch=cfg["mismatch"], gap_open=cfg["gap_open"], gap_ext=cfg["gap_ext"],
                           max_score=ms, read_size=rs, backtrace=cfg["backtrace"], reduce=cfg["reduce"], device=local_rank)
    bt = cfg["backtrace"]

    # synthetic pairs straight into pinned host memory (each rank its own index range of the stream)
    threads = max(1, (os.cpu_count() or 1) // world)
    h_plen, h_tlen = A.PinnedArray((P,), np.int32), A.PinnedArray((P,), np.int32)
    h_pat, h_txt = A.PinnedArray((P, rs), np.uint8), A.PinnedArray((P, rs), np.uint8)
    A.generate_pairs(cfg["seed"], P, cfg["length"], cfg["error"], rs, first_pair=shard.weak_first_pair(rank, P), nthreads=threads,
                     out=(h_plen.array, h_tlen.array, h_pat.array, h_txt.array))
    h_res = A.PinnedArray((P,), A.RESULT_DTYPE)
    h_ops = A.PinnedArray((P, 2 * rs), np.uint8) if bt else None

    # ---- device-resident arm ----
    d_plen = torch.from_numpy(h_plen.array).to(dev)
    d_tlen = torch.from_numpy(h_tlen.array).to(dev)
    d_pat = torch.from_numpy(h_pat.array).to(dev)
    d_txt = torch.from_numpy(h_txt.array).to(dev)
    d_res = torch.empty(P * A.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_ops = torch.empty((P, 2 * rs), dtype=torch.uint8, device=dev) if bt else None
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    launches = 0

    def device_step():
        nonlocal launches
        _, nl = A.align_device(params, P, d_plen.data_ptr(), d_tlen.data_ptr(), d_pat.data_ptr(), d_txt.data_ptr(),
                               d_res.data_ptr(), d_ops.data_ptr() if bt else None, stream=stream.cuda_stream,
                               device=local_rank, timed=False)
        launches += nl

    for _ in range(args.warmup):
        device_step()
    barrier()
    launches = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for k in range(args.steps):
        device_step()
        evs[k + 1].record(stream)
    barrier()
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    clocks = sampler.stop()
    total_ms_max = shard.max_over_ranks(total_ms, dev)
    ms_per_step = total_ms_max / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # sanity inside the bench: the timed path produced real alignments (scores in range, no failures)
    res_dev = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=A.RESULT_DTYPE)
    genasm = cfg["algo"].startswith("genasm")
    flagged = int((res_dev["status"] != 0).sum())
    # GenASM-DC: a small share of pairs has no defined reference output (status 3/4); nothing else may fail
    assert flagged == 0 or (genasm and flagged <= 0.02 * P), "bench: alignment failures on the timed path"
    mean_score = float(res_dev["score"].mean())
    assert (genasm and -1 <= mean_score) or 0 < mean_score <= ms + 1, "bench: implausible scores"

    # ---- end-to-end arm through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            A.align_batch(params, h_plen.array, h_tlen.array, h_pat.array, h_txt.array, results=h_res.array,
                          ops=h_ops.array if bt else None)
        e2e_step()  # warm the library's chunk buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        te_s = shard.max_over_ranks((time.perf_counter() - t0) / args.steps, dev)
        assert np.array_equal(h_res.array["score"], res_dev["score"]), "bench: e2e and device-resident scores differ"
        # op rows cross PCIe as run rows and are rebuilt into the caller's buffer by host threads inside the call (include/aim_b200.h)
        runs_pitch = 0
        if bt and os.environ.get("AIM_SPARSE_OPS", "1") != "0":
            runs_pitch = A.op_rows_download_bytes(params)
        e2e = {"value": world * P / te_s, "unit": "pairs/s",
               "h2d_bytes_per_step": int(P * (2 * rs + 8)),
               "d2h_bytes_per_step": int(P * (A.RESULT_DTYPE.itemsize + (runs_pitch if runs_pitch else (2 * rs if bt else 0)))),
               "ms_per_step": te_s * 1e3,
               "api": "aim_align_batch (C ABI), pinned host buffers, H2D+kernel+D2H double-buffered inside"}
        if runs_pitch:
            e2e["op_rows"] = {"download": (f"the first {runs_pitch} B of every row (its CIGAR string), copied to the head of the caller's {2 * rs}-byte rows by host "
                                           "threads inside the timed call (AIM_SPARSE_OPS=0: the rows themselves cross PCIe)") if genasm else
                                          (f"run rows of {runs_pitch} B per pair (op_runs_kernel), rebuilt into the caller's {2 * rs}-byte rows by host "
                                           "threads inside the timed call (AIM_SPARSE_OPS=0: the rows themselves cross PCIe)"),
                              "bytes_delivered_to_caller_per_step": int(P * (A.RESULT_DTYPE.itemsize + 2 * rs))}

    # ---- end-to-end arm with COMPACT transfers (aim_align_packed: 2-bit sequences in, run-length CIGAR rows out) ----
    e2e_packed = None
    if not args.no_e2e and cfg["algo"] == "wfa" and bt and rs < 2048:
        pitch = 64
        words = A.packed_row_bytes(rs) // 4
        h_packed = A.PinnedArray((P, 2, words), np.uint32)
        h_flags = A.PinnedArray(((P + 31) // 32,), np.uint32)
        h_cig = A.PinnedArray((P, pitch), np.uint8)
        h_res2 = A.PinnedArray((P,), A.RESULT_DTYPE)
        tp0 = time.perf_counter()
        A.pack_pairs(h_plen.array, h_tlen.array, h_pat.array, h_txt.array, rs, nthreads=threads, out=(h_packed.array, h_flags.array))
        pack_s = time.perf_counter() - tp0  # host-side, outside the timed region (the reference parses its pair file outside its timers too)

        def packed_step():
            A.align_packed(params, h_plen.array, h_tlen.array, h_packed.array, h_flags.array, cigar_pitch=pitch,
                           results=h_res2.array, cigars=h_cig.array)
        packed_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            packed_step()
        torch.cuda.synchronize(dev)
        tp_s = shard.max_over_ranks((time.perf_counter() - t0) / args.steps, dev)
        assert np.array_equal(h_res2.array["score"], res_dev["score"]), "bench: packed e2e and device-resident scores differ"
        assert int((h_res2.array["status"] != 0).sum()) == 0, "bench: packed e2e reported flagged / overflowing pairs"
        e2e_packed = {"value": world * P / tp_s, "unit": "pairs/s", "h2d_bytes_per_step": int(P * (2 * words * 4 + 8) + (P + 31) // 32 * 4),
                      "d2h_bytes_per_step": int(P * (A.RESULT_DTYPE.itemsize + pitch)), "ms_per_step": tp_s * 1e3,
                      "api": "aim_align_packed (C ABI extension): 2-bit packed sequences in, run-length CIGAR rows out, pinned host buffers",
                      "host_pack_pairs_per_s": P / pack_s, "host_pack_threads": threads,
                      "value_including_host_pack": world * P / (tp_s + shard.max_over_ranks(pack_s, dev)),
                      "note": "same scores and CIGAR text.  `value` times aim_align_packed on already packed buffers (a caller whose parser emits 2-bit "
                              "rows, as get_reads could); `value_including_host_pack` adds aim_pack_pairs on this rank's host threads, un-overlapped: "
                              "from the reference's ASCII buffers e2e_cigars (no host work) is the better route, from a pair FILE e2e_cli (parsed on the GPU)"}

    # ---- end-to-end arm with the reference's INPUT layout and CIGAR TEXT rows out (aim_align_batch_cigars) ----
    e2e_cigars = None
    if not args.no_e2e and bt and rs < 2048:
        cpitch = 32 if cfg["algo"] == "genasm_dc" and ms <= 8 else (64 if rs <= 200 else 96)
        h_cig2 = A.PinnedArray((P, cpitch), np.uint8)
        h_res3 = A.PinnedArray((P,), A.RESULT_DTYPE)

        def cig_step():
            A.align_batch_cigars(params, h_plen.array, h_tlen.array, h_pat.array, h_txt.array, cigar_pitch=cpitch,
                                 results=h_res3.array, cigars=h_cig2.array)
        cig_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cig_step()
        torch.cuda.synchronize(dev)
        tc_s = shard.max_over_ranks((time.perf_counter() - t0) / args.steps, dev)
        assert np.array_equal(h_res3.array["score"], res_dev["score"]), "bench: cigar-row e2e and device-resident scores differ"
        assert int((h_res3.array["status"] == 6).sum()) == 0, "bench: CIGAR rows overflowed"
        e2e_cigars = {"value": world * P / tc_s, "unit": "pairs/s", "h2d_bytes_per_step": int(P * (2 * rs + 8)),
                      "d2h_bytes_per_step": int(P * (A.RESULT_DTYPE.itemsize + cpitch)), "ms_per_step": tc_s * 1e3,
                      "api": "aim_align_batch_cigars (C ABI extension): the reference's input buffers, CIGAR text rows out, pinned host buffers"}

    # ---- end-to-end at the PROCESS boundary: pair file (page cache) -> output file, through aim_align_file (what `host <pairs> <out> <N>`
    # runs): file bytes up, parsed on the GPU, aligned, output text formatted on the GPU, text down, write() ----
    e2e_cli = None
    cli_out_path = None
    if not args.no_e2e and not genasm and not args.no_cli:
        shm = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
        cli_dir = Path(tempfile.mkdtemp(prefix=f"aimcli{rank}_", dir=shm))
        pairs_path, cli_out_path = cli_dir / "in.pairs", cli_dir / "out"
        A.write_pairs(pairs_path, h_plen.array, h_tlen.array, h_pat.array, h_txt.array)
        file_bytes = pairs_path.stat().st_size

        def cli_step():
            return A.align_file(params, pairs_path, cli_out_path, P, 1)
        cli_step()
        barrier()
        t0 = time.perf_counter()
        nst = max(2, min(args.steps, 5))
        for _ in range(nst):
            done, smask, cph, cl = cli_step()
        tcli_s = shard.max_over_ranks((time.perf_counter() - t0) / nst, dev)
        assert done == P and smask == 0, f"bench: aim_align_file aligned {done} of {P} pairs, status mask {smask}"
        out_bytes = cli_out_path.stat().st_size
        # the same call with the output discarded: what remains is reading, parsing, aligning and formatting (populating the pages of
        # a NEW output file is the operating system's file-write rate, about 4 GB/s on this box whatever the thread count)
        A.align_file(params, pairs_path, "/dev/null", P, 1)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            A.align_file(params, pairs_path, "/dev/null", P, 1)
        tnull_s = shard.max_over_ranks((time.perf_counter() - t0) / nst, dev)
        e2e_cli = {"value": world * P / tcli_s, "unit": "pairs/s", "ms_per_step": tcli_s * 1e3, "pair_file_bytes": file_bytes, "output_file_bytes": out_bytes,
                   "h2d_bytes_per_step": file_bytes, "d2h_bytes_per_step": out_bytes, "gpu_launches_per_step": cl,
                   "phase_ms_h2d_kernels_d2h": cph, "value_output_discarded": world * P / tnull_s,
                   "output_file": str(cli_out_path.parent.parent),
                   "api": "aim_align_file (C ABI; what `host <pairs> <out> <N>` calls): pair file in the page cache -> output file; get_reads and the "
                          "print loop run as GPU kernels, host threads only read() and write(); buffers allocated and freed inside the call"}
        pairs_path.unlink()

    # ---- parity: this step's results against the CPU oracle, bit-exact, OUTSIDE every timed region ----
    # Checked: the end-to-end arm's output (scores, spans and op bytes as they arrive in the caller's host buffers through
    # the C ABI); the device-resident arm's buffers must then be byte-identical to those.  Every rank checks its own pairs.
    parity = None
    if args.parity != "off":
        from oracle import oracle as O
        want = args.parity_pairs if args.parity_pairs >= 0 else (0 if args.parity == "full" else PARITY_SAMPLE[args.config])
        stride = 1 if want <= 0 or want >= P else max(1, P // want)
        tpar0 = time.perf_counter()
        if args.no_e2e:  # no host-side output yet: fetch the device-resident arm's
            h_res.array[:] = res_dev
            if bt:
                torch.from_numpy(h_ops.array).copy_(d_ops)
        kw = dict(max_score=ms, read_size=rs, mismatch=cfg["mismatch"], gap_open=cfg["gap_open"], gap_ext=cfg["gap_ext"],
                  backtrace=bt, reduce=cfg["reduce"], nthreads=threads, stride=stride, offset=(stride // 2))
        pr = O.check(cfg["algo"], h_plen.array, h_tlen.array, h_pat.array, h_txt.array, h_res.array, h_ops.array if bt else None, **kw)
        # device-resident arm == end-to-end arm, byte for byte (results incl. idx and status; op rows over the whole row)
        dev_same = bool(np.array_equal(res_dev.view(np.uint8), h_res.array.view(np.uint8)))
        if bt and dev_same:
            step_rows = max(1, (256 << 20) // (2 * rs))
            for a0 in range(0, P, step_rows):
                a1 = min(P, a0 + step_rows)
                if not torch.equal(d_ops[a0:a1], torch.from_numpy(h_ops.array[a0:a1]).to(dev, non_blocking=False)):
                    dev_same = False
                    break
        cig_checked = 0
        if e2e_cigars is not None and not genasm:  # CIGAR text rows (aim_align_batch_cigars / aim_align_packed) against the run-length text of the checked op rows
            samp = np.random.default_rng(1).choice(P, size=min(P, 200_000), replace=False)
            samp.sort()
            exp = A.cigar_strings(h_res.array[samp], h_ops.array[samp])
            for arr in [h_cig2.array] + ([h_cig.array] if e2e_packed is not None else []):
                got = [bytes(r).split(b"\0", 1)[0].decode() for r in arr[samp]]
                assert got == exp, "bench: CIGAR text rows differ from the run-length text of the op rows"
            if e2e_packed is not None:
                assert np.array_equal(h_res2.array["score"], h_res.array["score"]) and np.array_equal(h_res3.array["begin_offset"], h_res.array["begin_offset"])
            cig_checked = len(samp)
        cli_same = None
        if e2e_cli is not None and not args.no_e2e:  # the CLI path's output file == the host printer's text of the checked results
            ref_out = cli_out_path.with_name("expect")
            A.write_results(ref_out, h_res.array, h_ops.array if bt else None, rs, bt)
            cli_same = subprocess.run(["cmp", "-s", str(ref_out), str(cli_out_path)]).returncode == 0
            ref_out.unlink()
            assert cli_same, "bench: aim_align_file's output file differs from the printed text of the checked results"
        tot = torch.tensor([pr["pairs_checked"], pr["mismatches"], 0 if dev_same else 1], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        parity = {"pairs_checked": int(tot[0]), "mismatches": int(tot[1]), "device_arm_differs_on_ranks": int(tot[2]),
                  "against": "oracle/aim_oracle.c (C restatement pinned on the reference's outputs), score + status + begin/end offsets + op bytes of the span",
                  "checked_output": "aim_align_batch host buffers (e2e arm); device-resident buffers byte-compared to them",
                  "stride": stride, "cigar_text_rows_checked_per_rank": cig_checked, "cli_output_file_identical": cli_same, "oracle_threads_per_rank": threads,
                  "seconds": time.perf_counter() - tpar0, "first_bad_rank0": pr["first_bad"]}
        assert parity["mismatches"] == 0 and parity["device_arm_differs_on_ranks"] == 0, f"bench: PARITY FAILURE {parity}"

    # ---- one process, N GPUs: aim_align_batch(ngpus = N) on rank 0's batch while the other ranks wait (N > 1 only) ----
    inproc = None
    if world > 1 and not args.no_inproc and not args.no_e2e:
        gwait = dist.new_group(backend="gloo")
        barrier()
        if rank == 0:
            p_all = A.AlignParams(algo=cfg["algo"], mismatch=cfg["mismatch"], gap_open=cfg["gap_open"], gap_ext=cfg["gap_ext"], max_score=ms,
                                  read_size=rs, backtrace=bt, reduce=cfg["reduce"], device=0, ngpus=world)
            h_resN = A.PinnedArray((P,), A.RESULT_DTYPE)
            h_opsN = A.PinnedArray((P, 2 * rs), np.uint8) if bt else None

            def inproc_step():
                A.align_batch(p_all, h_plen.array, h_tlen.array, h_pat.array, h_txt.array, results=h_resN.array, ops=h_opsN.array if bt else None)
            inproc_step()
            t0 = time.perf_counter()
            nst = max(2, min(args.steps, 5))
            for _ in range(nst):
                inproc_step()
            ti_s = (time.perf_counter() - t0) / nst
            same = bool(np.array_equal(h_resN.array.view(np.uint8), h_res.array.view(np.uint8))) and (not bt or bool(np.array_equal(h_opsN.array, h_ops.array)))
            inproc = {"api": f"aim_align_batch(ngpus={world}) from ONE process: rank 0's {P} pairs split over {world} GPUs by index, pinned host buffers",
                      "value": P / ti_s, "unit": "pairs/s", "ms_per_step": ti_s * 1e3, "identical_to_one_gpu_output": same}
            assert same, "bench: aim_align_batch(ngpus=N) output differs from the one-GPU output"
            del h_resN, h_opsN
        dist.barrier(group=gwait)

    if rank == 0:
        # ---- rooflines for the dominant kernel (one launch = one step on one GPU) ----
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        kernel_s = statistics.median(step_ms) * 1e-3
        pl_mean, tl_mean = float(h_plen.array.mean()), float(h_tlen.array.mean())
        if genasm:
            # bytes: 2-bit sequences + lengths in, result + the CIGAR string out (runs of ~3 characters; measured from this run)
            cig_mean = float(res_dev["end_offset"].mean()) if bt else 0.0
            bytes_pair = (pl_mean / 4 + tl_mean / 4 + 8) + (8 + cig_mean)
            int_ops_pair = tl_mean * (ms + 1) * ((pl_mean + 64) // 64) * INT_OPS_PER_GENASM_WORD
        elif cfg["algo"] == "wfa":
            # bytes: 2-bit sequences + lengths in, result + 2-bit ops out
            bytes_pair = (pl_mean / 4 + tl_mean / 4 + 8) + (8 + ((pl_mean + tl_mean) / 4 if bt else 0))
            # 11 int-ops per computed (score, diagonal) cell (its I, D and M recurrences together) + one 4-op extend
            # group per live diagonal and score and per 16 matched bases
            sched = _wfa_work(res_dev["score"], cfg, ms)
            int_ops_pair = sched["diag_steps"] * INT_OPS_PER_OFFSET + (sched["diag_steps"] + pl_mean / 16) * INT_OPS_PER_EXTEND
            int_ops_pair_components = sched["offsets"] * INT_OPS_PER_OFFSET + (sched["diag_steps"] + pl_mean / 16) * INT_OPS_PER_EXTEND
        else:
            cells = pl_mean * tl_mean
            bytes_pair = (pl_mean / 4 + tl_mean / 4 + 8) + (8 + ((pl_mean + tl_mean) / 4 if bt else 0))
            int_ops_pair = cells * ((14 if bt else 10) if cfg["algo"] == "swg" else (8 if bt else 6))
        achieved_gbs = bytes_pair * P / kernel_s / 1e9
        roofline = {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "traffic": _ncu_traffic(args.config, P), "traffic_source": _ncu_entry(args.config).get("source"),
                    "peak_source": peak_src, "algorithmic_bytes_per_pair": bytes_pair,
                    "kernel": KERNELS[args.config], "kernel_ms": kernel_s * 1e3,
                    "note": "integer DP / bit-vector work: the binding ceiling is the INT32 ALU pipe (int_roofline), not HBM"}
        int_peak = A.measure_int_peak(local_rank)
        int_roofline = {"bound": "int32_alu", "achieved": int_ops_pair * P / kernel_s / 1e12, "peak": int_peak / 1e12, "unit": "Tops/s",
                        "frac": int_ops_pair * P / kernel_s / int_peak, "algorithmic_int_ops_per_pair": int_ops_pair,
                        "convention": "SURVEY 8d per-unit figures x units: WFA 11 ops per computed (score, diagonal) cell + 4 per extend group; "
                                      "NW 8 / SWG 14 ops per DP cell with backtrace flags",
                        "peak_source": "aim_measure_int_peak: dependent-free add/logic/min-max mix, this GPU, this run",
                        "gcups_equiv": pl_mean * tl_mean * P / kernel_s / 1e9}
        if cfg["algo"] == "wfa":
            int_roofline["frac_counting_each_component_offset"] = int_ops_pair_components * P / kernel_s / int_peak
            int_roofline["cells_per_pair"] = sched["diag_steps"]
            int_roofline["component_offsets_per_pair"] = sched["offsets"]
        cpu_baseline = None
        if not args.no_cpu_baseline:
            try:
                cthreads = os.cpu_count() or 1
                sample = args.ref_pairs or REF_SAMPLE[args.config]
                r = cpu_reference_run(cfg, sample, cthreads)
                tot = (r["h2d_ms"] + r["kernel_ms"] + r["d2h_ms"]) * 1e-3
                cpu_baseline = {"value": r["pairs"] / tot, "unit": "pairs/s", "cores": cthreads, "kind": "reference",
                                "sample": f"{r['pairs']} pairs of the same workload; reference DPU C sources built natively ({r['binary']}), "
                                          f"{r['nr_dpus']} simulated DPUs on {cthreads} host threads; CPU-DPU+DPU Kernel+DPU-CPU phases",
                                "kernel_only_value": r["pairs"] / (r["kernel_ms"] * 1e-3),
                                "upmem_functional_simulator": "unavailable (SDK not installed, no network)"}
            except Exception as ex:  # keep the GPU numbers even if the CPU leg cannot run
                cpu_baseline = {"value": None, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        line = {
            "metric": "aligned pairs/sec (score+CIGAR)", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64" if genasm else "int16", "data": "synthetic",
            "config": config_dict(cfg, P, world), "mean_score": mean_score, "parity": parity, "inproc_ngpus": inproc,
            "gcups_equiv": pl_mean * tl_mean * world * P / (ms_per_step * 1e-3) / 1e9,
            "clocks": clocks, "e2e": e2e, "e2e_packed": e2e_packed, "e2e_cigars": e2e_cigars, "e2e_cli": e2e_cli, "gpu_launches": launches,
            "roofline": roofline, "int_roofline": int_roofline, "cpu_baseline": cpu_baseline,
            "step_ms": step_ms,
        }
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if cli_out_path is not None:
        import shutil
        shutil.rmtree(cli_out_path.parent, ignore_errors=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    A.shutdown()


def _wfa_work(scores, cfg, max_score):
    """Mean computed offsets and extend probes per pair, from the data-independent wavefront schedule
    (widths per score) and the measured final scores of this run."""
    import numpy as np
    x, o, e = cfg["mismatch"], cfg["gap_open"], cfg["gap_ext"]
    pres = [True] + [False] * max_score
    lo, hi = [0] * (max_score + 1), [0] * (max_score + 1)
    hasI, hasD = [False] * (max_score + 1), [False] * (max_score + 1)
    cum_off, cum_diag = [0.0] * (max_score + 2), [0.0] * (max_score + 2)
    cum_off[0] = cum_diag[0] = 1.0
    for s in range(1, max_score + 1):
        A_ = s - x >= 0 and pres[s - x]
        B_ = s - o - e >= 0 and pres[s - o - e]
        E_ = s - e >= 0 and pres[s - e]
        ie = E_ and hasI[s - e]
        de = E_ and hasD[s - e]
        io, do = B_ or ie, B_ or de
        if A_ or io or do:
            los = [lo[s - x]] if A_ else []
            his = [hi[s - x]] if A_ else []
            if B_:
                los.append(lo[s - o - e]); his.append(hi[s - o - e])
            if ie or de:
                los.append(lo[s - e]); his.append(hi[s - e])
            pres[s], lo[s], hi[s], hasI[s], hasD[s] = True, min(los) - 1, max(his) + 1, io, do
            w = hi[s] - lo[s] + 1
            cum_off[s], cum_diag[s] = cum_off[s - 1] + w * (1 + io + do), cum_diag[s - 1] + w
        else:
            cum_off[s], cum_diag[s] = cum_off[s - 1], cum_diag[s - 1]
    cum_off[max_score + 1], cum_diag[max_score + 1] = cum_off[max_score], cum_diag[max_score]
    sc = np.clip(scores, 0, max_score + 1)
    if cfg["length"] > 2000:  # adaptive trimming is active on long reads: widths are not schedule widths
        return {"offsets": float(np.mean(sc)) * 139.0, "diag_steps": float(np.mean(sc)) * 46.0}
    return {"offsets": float(np.mean(np.asarray(cum_off)[sc])), "diag_steps": float(np.mean(np.asarray(cum_diag)[sc]))}


def _ncu_entry(config: int) -> dict:
    """The committed `ncu --set full` summary of this config's dominant kernel (profiles/ncu_summary.json)."""
    f = ROOT / "profiles" / "ncu_summary.json"
    try:
        return json.loads(f.read_text()).get(str(config), {})
    except Exception:
        return {}


def _ncu_traffic(config: int, pairs: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel for one launch over `pairs`
    pairs: measured per pair in the committed capture, scaled to this launch (None without a capture)."""
    per_pair = _ncu_entry(config).get("dram_bytes_per_pair")
    return None if per_pair is None else per_pair * pairs


if __name__ == "__main__":
    main()
