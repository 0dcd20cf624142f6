#!/bin/bash
# second GPU call: large-sample ablation with the fp32-reorder floor, precision profiles through tests and bench
set -x
mkdir -p gpurun_out/j2
O=gpurun_out/j2
timeout 900 python profiles/experiments/pass_ablation.py --workload c3 --total 512 --chunk 64 --set mixes > $O/pass_ablation_c3_mixes.md 2> $O/pass_ablation_c3_mixes.err
timeout 600 python profiles/experiments/pass_ablation.py --workload c2 --total 512 --chunk 32 --set mixes > $O/pass_ablation_c2_mixes.md 2> $O/pass_ablation_c2_mixes.err
for P in 1 2; do
  EIG_PRECISION=$P timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > $O/pytest_gpu_precision$P.txt
done
for P in 0 1 2; do
  EIG_PRECISION=$P timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_c3_precision$P.json 2> $O/bench_c3_precision$P.err
done
ls -la $O
