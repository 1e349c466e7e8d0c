#!/bin/bash
# dp_scan_kernel on the GPU box: parity tests of every mode, then an A/B of the modes on configs 3 and 2 (device-resident,
# every pair checked against the oracle), then the launch list of config 3.  usage: tools/iter_scan.sh [tag]
tag=${1:-scan}
CFG3_SET=${CFG3_SET:-"AIM_DP_SCAN=0 AIM_DP_SCAN=1 AIM_DP_SCAN=2"}
CFG2_SET=${CFG2_SET:-"AIM_DP_SCAN=0 AIM_DP_SCAN=1 AIM_DP_SCAN=2"}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_kernel" > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d.get('parity') or {}; print('%.3e pairs/s %.2f ms parity %s/%s dev_differs %s' % (d['value'], d['ms_per_step'], p.get('mismatches'), p.get('pairs_checked'), p.get('device_arm_differs_on_ranks')))"; }
for cfg in 3 2; do
  if [ $cfg = 3 ]; then set -- $CFG3_SET; else set -- $CFG2_SET; fi
  for kv in "$@"; do
    out=gpurun_out/${tag}_cfg${cfg}_$(echo $kv | tr ' =' '__').json
    env $(echo $kv | tr "," " ") timeout 200 python bench.py --config $cfg --no-cpu-baseline --no-e2e --no-cli --parity-pairs 200000 --steps 5 --warmup 3 2>gpurun_out/${tag}_err.log | tail -1 > $out
    echo "cfg $cfg $kv  $(line < $out)"
  done
done
for kv in "AIM_DP_SCAN=1"; do
  env $kv timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${tag}_launches_cfg3_$(echo $kv | tr ' =' '__').csv \
    python bench.py --config 3 --no-cpu-baseline --no-e2e --no-cli --parity off --steps 1 --warmup 1 > /dev/null 2>&1
done
python tools/launch_table.py gpurun_out/${tag}_launches_cfg3_AIM_DP_SCAN_1.csv 2>/dev/null | tail -12
if [ -n "$NO_FULL" ]; then exit 0; fi
# one --set full capture of the fill kernel (200 K pairs)
AIM_DP_SCAN=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dp_scan_kernel -c 1 -f -o gpurun_out/${tag}_dp_scan_cfg3 \
    python bench.py --config 3 --steps 1 --warmup 1 --pairs 200000 --no-cpu-baseline --no-e2e --no-cli --parity off > /dev/null 2>&1
ls -la gpurun_out/${tag}_dp_scan_cfg3.ncu-rep
