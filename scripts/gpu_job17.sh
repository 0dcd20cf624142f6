#!/bin/bash
set -x
mkdir -p gpurun_out/j17
O=gpurun_out/j17
timeout 600 python bench.py --steps 20 --warmup 5 --also '' --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt
