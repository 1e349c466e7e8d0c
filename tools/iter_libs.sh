#!/bin/bash
# A/B of prebuilt library variants on config 4 (device-resident): usage tools/iter_libs.sh build/variants/a.so build/variants/b.so ...
# (each variant replaces aim_b200/libaim_b200.so on the GPU box for its run; the first also runs the WFA parity tests)
first=1
for lib in "$@"; do
  cp "$lib" aim_b200/libaim_b200.so
  if [ $first = 1 ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wfa or WFA" 2>&1 | tail -2; first=0; fi
  v=$(timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1fM pairs/s %.2f ms' % (d['value']/1e6, d['ms_per_step']))")
  echo "$lib  $v"
done
