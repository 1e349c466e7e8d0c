#!/bin/bash
# round-2 closing evidence in one GPU call (gpurun -- 'bash tools/final_run2.sh r02z'): full -m gpu suite, smoke, the bench line of
# config 4 with the op rows as run rows (default) and as they are (AIM_SPARSE_OPS=0) on the SAME box, the reference arm, configs
# 2 / 3 / 5, the launch list of a config-4 step, an ncu capture of op_runs_kernel on the e2e path, and compute-sanitizer memcheck of
# the run-row tests.
tag=${1:-r02z}
out=gpurun_out
mkdir -p $out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3) > $out/${tag}_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> $out/${tag}_tests.log
timeout 300 python bench.py > $out/${tag}_bench_cfg4.json 2> $out/${tag}_bench_cfg4.err
AIM_SPARSE_OPS=0 timeout 300 python bench.py --no-cli --no-cpu-baseline --parity off > $out/${tag}_bench_cfg4_direct_rows.json 2>> $out/${tag}_bench_cfg4.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref_cfg4.json 2>/dev/null
for c in 2 3 5 6 7; do timeout 300 python bench.py --config $c > $out/${tag}_bench_cfg$c.json 2> $out/${tag}_bench_cfg$c.err; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_cfg4.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-cli --parity off > /dev/null 2>&1
cat > /tmp/e2e_small.py <<'PY'
import aim_b200 as A
ms, rs = A.derive_knobs("wfa", 150, 0.04)
arrs = A.generate_pairs(4, 600_000, 150, 0.04, rs)
p = A.AlignParams(algo="wfa", max_score=ms, read_size=rs, backtrace=True, reduce=True)
for _ in range(2):
    A.align_batch(p, *arrs)
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:op_runs_kernel -c 1 -f -o $out/${tag}_op_runs_cfg4 python /tmp/e2e_small.py > /dev/null 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_op_runs.py -m gpu -x -q -k "wfa or nw-24" > $out/${tag}_sanitizer_memcheck_op_runs.log 2>&1; echo "memcheck rc=$?" >> $out/${tag}_sanitizer_memcheck_op_runs.log
cat $out/${tag}_tests.log
tail -2 $out/${tag}_sanitizer_memcheck_op_runs.log
python tools/benchline.py $out/${tag}_bench_cfg*.json
