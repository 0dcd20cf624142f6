#!/bin/bash
set -x
mkdir -p gpurun_out/j11
O=gpurun_out/j11
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|full population|fold |precision |r_c5|GPU vs|assert|tc_check" > $O/pytest_gpu_all.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
EIG_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/kernel_metrics_c2.csv \
     python profiles/experiments/one_eval.py --workload c2 --evals 2 > $O/one_eval_c2.log 2>&1
gzip -f $O/kernel_metrics_c2.csv
ls -la $O
