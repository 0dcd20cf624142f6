#!/bin/bash
set -x
mkdir -p gpurun_out/j10
O=gpurun_out/j10
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|Error|error|full population|fold |precision |r_c5|GPU vs|assert|tc_check" > $O/pytest_gpu_all.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --workload c2 --steps 100 --warmup 5 --also '' > $O/bench_c2_main.json 2> $O/bench_c2_main.err
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck_smoke.txt 2>&1
EIG_FOLD=1 timeout 600 compute-sanitizer --tool memcheck python profiles/experiments/one_eval.py --workload c2 --pop 4 --evals 1 > $O/sanitizer_memcheck_fold.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck python profiles/experiments/racecheck_simt.py > $O/sanitizer_racecheck_simt.txt 2>&1
tail -5 $O/sanitizer_*.txt
ls -la $O
