#!/bin/bash
# A/B of env knobs on config 4 (device-resident): usage tools/iter_ab.sh "VAR=a VAR2=b" "VAR=c" ...
for kv in "$@"; do
  v=$(env $kv timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1fM pairs/s %.2f ms' % (d['value']/1e6, d['ms_per_step']))")
  echo "$kv  $v"
done
