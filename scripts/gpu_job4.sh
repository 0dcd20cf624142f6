#!/bin/bash
set -x
mkdir -p gpurun_out/j4
O=gpurun_out/j4
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|full population|fold |precision |r_c5|GPU vs|assert" > $O/pytest_gpu_all.txt
for F in 0 1; do
  EIG_FOLD=$F timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_fold$F.json 2> $O/bench_c3_fold$F.err
done
EIG_FOLD=1 EIG_PRECISION=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also '' > $O/bench_c3_fold1_precision1.json 2> $O/bench_c3_fold1_precision1.err
EIG_FOLD=1 EIG_PRECISION=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also '' > $O/bench_c3_fold1_precision2.json 2> $O/bench_c3_fold1_precision2.err
timeout 900 python profiles/experiments/pass_ablation.py --workload c3 --total 512 --chunk 64 --set layer3 > $O/pass_ablation_c3_layer3.md 2> $O/pass_ablation_c3_layer3.err
timeout 600 tests/gpu/tc_check time c3 7 > $O/tc_time_c3_passes7.log 2>&1
timeout 600 tests/gpu/tc_check time c3 4 > $O/tc_time_c3_passes4.log 2>&1
ls -la $O
