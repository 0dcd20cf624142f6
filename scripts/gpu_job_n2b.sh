#!/bin/bash
set -x
mkdir -p gpurun_out/n2b
O=gpurun_out/n2b
nproc > $O/nproc.txt; taskset -p $$ >> $O/nproc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --also '' > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err
