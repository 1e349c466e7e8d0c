#!/bin/bash
# one GPU call: config-4 bench line (full), ncu launch list of the same command, one --set full capture of wfa_sub_kernel
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wfa_sub_kernel -c 1 -f -o gpurun_out/${tag}_wfa_sub4_cfg4 \
    python bench.py --steps 1 --warmup 1 --pairs 2000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
tail -c 600 gpurun_out/${tag}_bench_cfg4.json
