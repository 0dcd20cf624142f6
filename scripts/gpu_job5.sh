#!/bin/bash
set -x
mkdir -p gpurun_out/j5
O=gpurun_out/j5
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
for F in 0 1; do
  EIG_FOLD=$F EIG_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/kernel_metrics_c3_fold$F.csv \
     python profiles/experiments/one_eval.py --workload c3 --evals 1 > $O/one_eval_c3_fold$F.log 2>&1
  gzip -f $O/kernel_metrics_c3_fold$F.csv
done
ls -la $O
