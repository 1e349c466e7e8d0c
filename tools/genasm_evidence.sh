tag=r01o
for c in 7 8 9; do timeout 300 python bench.py --config $c > gpurun_out/${tag}_bench_cfg$c.json 2> gpurun_out/${tag}_bench_cfg$c.err; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg7.csv \
    python bench.py --config 7 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:genasm -c 2 -f -o gpurun_out/${tag}_genasm_cfg7 \
    python bench.py --config 7 --steps 1 --warmup 1 --pairs 2000000 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python tools/benchline.py gpurun_out/${tag}_bench_cfg*.json
