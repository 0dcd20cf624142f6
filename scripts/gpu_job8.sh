#!/bin/bash
set -x
mkdir -p gpurun_out/j8
O=gpurun_out/j8
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|full population|fold |precision |r_c5|GPU vs|assert|tc_check" > $O/pytest_gpu_all.txt
for F in 0 1; do
  EIG_FOLD=$F timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3_fold$F.json 2> $O/bench_c3_fold$F.err
done
EIG_FOLD=1 EIG_PRECISION=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also '' > $O/bench_c3_fold1_precision1.json 2> $O/bench_c3_fold1_precision1.err
EIG_FOLD=1 EIG_PRECISION=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --also '' > $O/bench_c3_fold1_precision2.json 2> $O/bench_c3_fold1_precision2.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
EIG_FOLD=1 EIG_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/kernel_metrics_c3_fold1.csv \
     python profiles/experiments/one_eval.py --workload c3 --evals 1 > $O/one_eval_c3_fold1.log 2>&1
gzip -f $O/kernel_metrics_c3_fold1.csv
ls -la $O
