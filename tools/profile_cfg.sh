#!/bin/bash
# one GPU call for any bench config: bench line, ncu launch list of the same command, one --set full capture of the named kernel
#   tools/profile_cfg.sh <tag> <config> <kernel-regex> <capture-name> [pairs-in-capture]
tag=$1; cfg=$2; kre=$3; cap=$4; pairs=${5:-2000000}
mkdir -p gpurun_out
timeout 900 python bench.py --config $cfg > gpurun_out/${tag}_bench_cfg${cfg}.json 2> gpurun_out/${tag}_bench_cfg${cfg}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg${cfg}.csv \
    python bench.py --config $cfg --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -c 1 -f -o gpurun_out/${tag}_${cap} \
    python bench.py --config $cfg --steps 1 --warmup 1 --pairs $pairs --no-cpu-baseline --no-e2e > /dev/null 2>&1
tail -c 900 gpurun_out/${tag}_bench_cfg${cfg}.json
