#!/bin/bash
set -x
mkdir -p gpurun_out/n8
O=gpurun_out/n8
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err
tail -5 $O/bench_n8.err
ls -la $O
