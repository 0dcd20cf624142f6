#!/bin/bash
# Builds libeig.so for sm_100a in-tree (the .so is git-ignored but travels to the GPU box).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false \
      -Xcompiler -fPIC -shared ${EIG_NVCC_EXTRA:-} -o ../libeig.so eig_api.cu -lcudart_static -ldl -lrt -lpthread
# GPU self-check of the tcgen05 convolution (run by tests/test_gpu_parity.py on the B200)
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false ${EIG_NVCC_EXTRA:-} \
      -o ../../tests/gpu/tc_check ../../tests/gpu/tc_check.cu -lcudart_static -ldl -lrt -lpthread
# host glue: the genome flattener as a CPython extension (genome.py: flatten_genome_fast)
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
gcc -O2 -fPIC -shared -Wall -I"$PYINC" -o ../_flatten.so flatten.c
echo "built $(cd .. && pwd)/libeig.so and _flatten.so"
