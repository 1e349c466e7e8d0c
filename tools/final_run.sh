#!/bin/bash
# round-end evidence in one GPU call: full -m gpu suite, smoke, bench lines of every config, the reference arm, launch lists and
# --set full captures for config 4 (wfa_sub_kernel) and config 7 (genasm band + traceback kernels)
tag=${1:-r01n}
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > gpurun_out/${tag}_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> gpurun_out/${tag}_tests.log
timeout 300 python bench.py > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref_cfg4.json 2>/dev/null
for c in 2 3 5 6 7 8 9; do timeout 300 python bench.py --config $c > gpurun_out/${tag}_bench_cfg$c.json 2> gpurun_out/${tag}_bench_cfg$c.err; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg7.csv \
    python bench.py --config 7 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wfa_sub_kernel -c 1 -f -o gpurun_out/${tag}_wfa_sub4_cfg4 \
    python bench.py --steps 1 --warmup 1 --pairs 2000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:genasm -c 2 -f -o gpurun_out/${tag}_genasm_cfg7 \
    python bench.py --config 7 --steps 1 --warmup 1 --pairs 2000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/${tag}_tests.log
python tools/benchline.py gpurun_out/${tag}_bench_cfg*.json
