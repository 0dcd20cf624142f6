#!/bin/bash
# third GPU call: fold on/off, tightened parity gates, per-role cycles of the C3 shapes at 3 products and 1 product
set -x
mkdir -p gpurun_out/j3
O=gpurun_out/j3
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -120 > $O/pytest_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|full population|fold |precision |whole path|r_c5|GPU vs" > $O/pytest_gpu_all.txt
for F in 0 1; do
  EIG_FOLD=$F timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_fold$F.json 2> $O/bench_c3_fold$F.err
done
EIG_FOLD=1 EIG_PRECISION=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also '' > $O/bench_c3_fold1_precision1.json 2> $O/bench_c3_fold1_precision1.err
timeout 600 tests/gpu/tc_check time c3 7 > $O/tc_time_c3_passes7.log 2>&1
timeout 600 tests/gpu/tc_check time c3 4 > $O/tc_time_c3_passes4.log 2>&1
ls -la $O
