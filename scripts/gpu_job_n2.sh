#!/bin/bash
# 2-GPU call: NCCL parity test + the contract bench at N=2 (torchrun, as the driver launches it)
set -x
mkdir -p gpurun_out/n2
O=gpurun_out/n2
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "two_rank" 2>&1 | tail -15 > $O/pytest_nccl.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -5 $O/bench_n2.err
ls -la $O
