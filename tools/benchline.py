#!/usr/bin/env python
"""One-line digest of a bench.py JSON line:  python tools/benchline.py <file.json> [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "unreadable:", ex)
        continue
    e2e = d.get("e2e") or {}
    cpu = d.get("cpu_baseline") or {}
    pk = d.get("e2e_packed") or {}
    cg = d.get("e2e_cigars") or {}
    print("%s | value %.2fM e2e %.2fM e2e_cigars %.2fM e2e_packed %.2fM cpu %.3fM | %.2f ms/step | launches %s | int frac %.3f" % (
        d["config"]["workload"][:48], d["value"] / 1e6, (e2e.get("value") or 0) / 1e6, (cg.get("value") or 0) / 1e6, (pk.get("value") or 0) / 1e6, (cpu.get("value") or 0) / 1e6,
        d["ms_per_step"], d.get("gpu_launches"), (d.get("int_roofline") or {}).get("frac", 0)))
