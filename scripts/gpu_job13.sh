#!/bin/bash
set -x
mkdir -p gpurun_out/j13
O=gpurun_out/j13
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
tail -2 $O/smoke.txt
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
for wl in c2 c3; do
EIG_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/kernel_metrics_$wl.csv python profiles/experiments/one_eval.py --workload $wl --evals 1 > $O/one_eval_$wl.log 2>&1
gzip -f $O/kernel_metrics_$wl.csv
done
ls -la $O
