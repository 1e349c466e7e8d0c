#!/bin/bash
# Evidence pack of a round, run on the GPU box (gpurun -- 'bash tools/profile_round.sh r02j'):
#   launch lists (ncu gpu__time_duration, one step) and one `ncu --set full --import-source on` capture of the dominant kernel of
#   every BASELINE config, plus compute-sanitizer memcheck / racecheck logs of the parity tests on small inputs.
# The .ncu-rep files come back in gpurun_out/; tools/ncu_summary.py turns them into profiles/<tag>_*.{txt,json}.
tag=${1:-r02}
out=gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cli --no-e2e --parity off"
for c in 2 3 4 5 6; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_cfg$c.csv $B --config $c > /dev/null 2>&1
done
full() {  # cfg kernel-regex name pairs
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -o $out/${tag}_$3_cfg$1 $B --config $1 --pairs $4 > /dev/null 2>&1
}
full 4 wfa_sub_kernel wfa_sub 2000000
full 3 dp2_strip_kernel dp2_strip 200000
full 3 dp_scan_kernel dp_scan 200000
full 2 dp2_strip_kernel dp2_strip 400000
full 5 wfa_long_kernel wfa_long 20000
full 6 wfa_long_kernel wfa_long_bt 8000
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_file.py tests/test_packed.py -m gpu -x -q -k "ragged or golden or arena or chunks or adversarial or packed" > $out/${tag}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/${tag}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged and (acgt) and not literal" > $out/${tag}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $out/${tag}_sanitizer_racecheck.log
tail -3 $out/${tag}_sanitizer_memcheck.log $out/${tag}_sanitizer_racecheck.log
ls -la $out/${tag}_*
