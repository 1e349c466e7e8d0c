#!/bin/bash
# one short GPU call for the fused traceback of dp_scan_kernel: the NW/SWG-related GPU tests, config 3 with the traceback inside the
# fill kernel (default) and as a kernel of its own on the same box, the launch list.
tag=${1:-r02R}
out=gpurun_out
mkdir -p $out
(timeout 150 python -m pytest tests -m gpu -x -q -k "scan_kernel or nw or swg or NW or SWG or golden or cfg3 or cfg2 or dp_" 2>&1 | tail -3) > $out/${tag}_tests_dp.log
cat $out/${tag}_tests_dp.log
timeout 120 python bench.py --config 3 --no-cpu-baseline --no-cli > $out/${tag}_bench_cfg3.json 2> $out/${tag}_bench_cfg3.err
AIM_DP_SCAN_TB=kernel timeout 100 python bench.py --config 3 --no-cpu-baseline --no-cli --no-e2e --parity-pairs 200000 > $out/${tag}_bench_cfg3_tb_kernel.json 2>> $out/${tag}_bench_cfg3.err
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $out/${tag}_launches_cfg3.csv \
    python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-cli --parity off > /dev/null 2>&1
python tools/benchline.py $out/${tag}_bench_cfg3*.json
python -c "
import json
for f in ('$out/${tag}_bench_cfg3.json', '$out/${tag}_bench_cfg3_tb_kernel.json'):
    d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['parity']['pairs_checked'], d['parity']['mismatches'], d['parity']['device_arm_differs_on_ranks'], d['gpu_launches'])
"
python tools/launch_table.py $out/${tag}_launches_cfg3.csv | cut -c1-110
