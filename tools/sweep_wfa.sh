#!/bin/bash
# sweep the lockstep WFA kernel's launch knobs on config 4 (device-resident only)
for g in 4 8 16; do for w in 1 2 4; do
  v=$(AIM_WFA_G=$g AIM_WFA_WPB=$w timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 3 --warmup 2 --pairs 4000000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1fM pairs/s %.2f ms' % (d['value']/1e6, d['ms_per_step']))")
  echo "G=$g WPB=$w  $v"
done; done
