#!/bin/bash
# first GPU call of round 2: sanity tests, baseline bench lines, pass ablation, per-kernel dram bytes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3_base.json 2> gpurun_out/bench_c3_base.err
timeout 300 python bench.py --workload c2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_base.json 2> gpurun_out/bench_c2_base.err
timeout 600 python profiles/experiments/pass_ablation.py --workload c3 --pop 64 > gpurun_out/pass_ablation_c3.md 2> gpurun_out/pass_ablation_c3.err
timeout 600 python profiles/experiments/pass_ablation.py --workload c2 --pop 32 > gpurun_out/pass_ablation_c2.md 2> gpurun_out/pass_ablation_c2.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
for wl in c2 c3; do
  EIG_NO_GRAPH=1 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/kernel_metrics_$wl.csv \
     python profiles/experiments/one_eval.py --workload $wl --evals 2 > gpurun_out/one_eval_$wl.log 2>&1
  gzip -f gpurun_out/kernel_metrics_$wl.csv
done
ls -la gpurun_out
