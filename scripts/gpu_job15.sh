#!/bin/bash
set -x
mkdir -p gpurun_out/j15
O=gpurun_out/j15
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|full population|fold |precision |r_c5|GPU vs|assert|tc_check" > $O/pytest_gpu_all.txt
tail -3 $O/pytest_gpu_all.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
ls -la $O
