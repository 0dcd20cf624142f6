#!/bin/bash
set -x
mkdir -p gpurun_out/j12
O=gpurun_out/j12
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|full population|fold |precision |r_c5|GPU vs|assert|tc_check" > $O/pytest_gpu_all.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
for P in 1 2; do
EIG_PRECISION=$P timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_precision$P.json 2> $O/bench_c3_precision$P.err
done
EIG_NO_GRAPH=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:l0_conva1_tile -s 10 -c 1 -o $O/a1tile_c2 python profiles/experiments/one_eval.py --workload c2 --evals 1 > $O/ncu_a1.log 2>&1
timeout 300 tests/gpu/tc_check time 2>&1 | cut -c1-112 > $O/tc_time_c2.log
timeout 300 tests/gpu/tc_check time c3 4 2>&1 | cut -c1-112 > $O/tc_time_c3_p4.log
timeout 300 tests/gpu/tc_check time c3 7 2>&1 | cut -c1-112 > $O/tc_time_c3_p7.log
ls -la $O
