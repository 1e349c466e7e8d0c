#!/usr/bin/env python
"""Summarise an `ncu --set full` capture into profiles/ (tracked): key counters + per-source-line shares.

    python tools/ncu_summary.py <report.ncu-rep> <kernel-substring> <out-prefix> [pairs-in-launch] [cubin] [mangled-kernel-substring]

The mangled substring (e.g. wfa_sub_kernelILi4ELb1ELb1) must select ONE template instantiation in the cubin: the join in
ncu_lines.py is by instruction order.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
]
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, kern, prefix = sys.argv[1:4]
    pairs = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    cubin = sys.argv[5] if len(sys.argv) > 5 else None
    mangled = sys.argv[6] if len(sys.argv) > 6 else kern
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    sel = [r for r in rows[2:] if kern in r[ki]]
    out = []
    summary = {"report": Path(rep).name, "kernel": kern, "launches_captured": len(sel), "pairs_in_launch": pairs}
    for n, r in enumerate(sel):
        d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        out.append(f"== launch {n}: {r[ki]}")
        for k in KEYS:
            if k in d:
                out.append(f"{k:90s} {d[k][0]:>18s} {d[k][1]}")
        if n == 0:
            rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * UNIT.get(d["dram__bytes_read.sum"][1], 1.0)
            wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * UNIT.get(d["dram__bytes_write.sum"][1], 1.0)
            summary.update(dram_bytes_per_launch=rd + wr, dram_read_bytes=rd, dram_write_bytes=wr,
                           duration_ms=float(d["gpu__time_duration.sum"][0].replace(",", "")) * (1e-3 if d["gpu__time_duration.sum"][1] == "us" else 1.0),
                           alu_pipe_pct=float(d["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][0]),
                           issue_active_pct=float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                           warp_inst=float(d["smsp__inst_executed.sum"][0].replace(",", "")))
            if pairs:
                summary["dram_bytes_per_pair"] = (rd + wr) / pairs
                summary["warp_inst_per_pair"] = summary["warp_inst"] / pairs
    if cubin:
        here = Path(__file__).resolve().parent
        lines = subprocess.run([sys.executable, str(here / "ncu_lines.py"), rep, cubin, mangled, "40"],
                               capture_output=True, text=True).stdout
        out.append("\n== instruction share per CUDA source line (top 40) ==\n" + lines)
    Path(prefix + ".txt").write_text("\n".join(out) + "\n")
    Path(prefix + ".json").write_text(json.dumps(summary, indent=1) + "\n")
    print("\n".join(out[:45]))


if __name__ == "__main__":
    main()
