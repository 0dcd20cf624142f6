#!/bin/bash
set -x
mkdir -p gpurun_out/n4
O=gpurun_out/n4
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_n4.json 2> $O/bench_n4.err
tail -3 $O/bench_n4.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > $O/bench_ref_n4.json 2> $O/bench_ref_n4.err
ls -la $O
