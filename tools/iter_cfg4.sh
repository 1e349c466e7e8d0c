#!/bin/bash
# one GPU iteration on the short-read WFA kernel: WFA parity tests, then the config-4 device-resident number
tag=${1:-it}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "wfa or WFA" 2>&1 | tail -4) > gpurun_out/${tag}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_tests.log
python - gpurun_out/${tag}_bench.json <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cfg4 %.1fM pairs/s  %.2f ms" % (d["value"]/1e6, d["ms_per_step"]))
P
