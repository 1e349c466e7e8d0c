#!/bin/bash
# A/B of prebuilt library variants (build/variants/lib_*.so) x env settings on config 3, device-resident, 200 K pairs checked:
#   tools/iter_scan_libs.sh <tag> "<lib ...>" "<env-set ...>"   (an env set = VAR=a,VAR2=b)
tag=$1; libs=$2; sets=$3
mkdir -p gpurun_out
cp aim_b200/libaim_b200.so /tmp/lib_keep.so
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d.get('parity') or {}; print('%.3e pairs/s %.2f ms parity %s/%s dev_differs %s' % (d['value'], d['ms_per_step'], p.get('mismatches'), p.get('pairs_checked'), p.get('device_arm_differs_on_ranks')))"; }
for lib in $libs; do
  cp build/variants/lib_$lib.so aim_b200/libaim_b200.so
  for kv in $sets; do
    out=gpurun_out/${tag}_cfg3_${lib}_$(echo $kv | tr ',=' '__').json
    env $(echo $kv | tr "," " ") timeout 200 python bench.py --config 3 --no-cpu-baseline --no-e2e --no-cli --parity-pairs 200000 --steps 5 --warmup 3 2>gpurun_out/${tag}_err.log | tail -1 > $out
    echo "lib $lib $kv  $(line < $out)"
  done
done
cp /tmp/lib_keep.so aim_b200/libaim_b200.so
