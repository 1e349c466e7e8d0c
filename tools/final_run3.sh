#!/bin/bash
# closing evidence after dp_scan_kernel (gpurun -- 'bash tools/final_run3.sh r02N'): full -m gpu suite, smoke, the bench lines of
# configs 3, 2 and 4, config 3 with the scan kernel off on the SAME box, the launch list of a config-3 step, one --set full
# capture of dp_scan_kernel, compute-sanitizer memcheck of the scan-kernel tests.
tag=${1:-r02N}
out=gpurun_out
mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $out/${tag}_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> $out/${tag}_tests.log
cat $out/${tag}_tests.log
timeout 300 python bench.py --config 3 > $out/${tag}_bench_cfg3.json 2> $out/${tag}_bench_cfg3.err
AIM_DP_SCAN=0 timeout 300 python bench.py --config 3 --no-cli --no-cpu-baseline --parity-pairs 200000 > $out/${tag}_bench_cfg3_scan_off.json 2>> $out/${tag}_bench_cfg3.err
if [ -z "$ONLY_CFG3" ]; then
timeout 300 python bench.py --config 2 > $out/${tag}_bench_cfg2.json 2> $out/${tag}_bench_cfg2.err
timeout 300 python bench.py > $out/${tag}_bench_cfg4.json 2> $out/${tag}_bench_cfg4.err
fi
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_cfg3.csv \
    python bench.py --config 3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-cli --parity off > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dp_scan_kernel -c 1 -f -o $out/${tag}_dp_scan_cfg3 \
    python bench.py --config 3 --steps 1 --warmup 1 --pairs 200000 --no-cpu-baseline --no-e2e --no-cli --parity off > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scan_kernel and (swg-24 or nw-27 or swg-15) and (1-0 or 2-0)" > $out/${tag}_sanitizer_memcheck_dp_scan.log 2>&1; echo "memcheck rc=$?" >> $out/${tag}_sanitizer_memcheck_dp_scan.log
tail -3 $out/${tag}_sanitizer_memcheck_dp_scan.log
python tools/benchline.py $out/${tag}_bench_cfg*.json
