#!/bin/bash
set -x
mkdir -p gpurun_out/j18
O=gpurun_out/j18
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err; tail -2 $O/bench_c3.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 > $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
