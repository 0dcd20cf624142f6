#!/bin/bash
set -x
mkdir -p gpurun_out/j9
O=gpurun_out/j9
timeout 900 python bench.py --steps 10 --warmup 3 --also c2,c4,c5 > $O/bench_c3_also_all.json 2> $O/bench_c3_also_all.err
tail -3 $O/bench_c3_also_all.err
for F in 0 1; do
  EIG_FOLD=$F timeout 300 python bench.py --steps 20 --warmup 3 --pop 16 --also '' --no-cpu-baseline --parity 0 > $O/bench_c3_pop16_fold$F.json 2> $O/bench_c3_pop16_fold$F.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_c3.json 2> $O/bench_reference_c3.err
ls -la $O
