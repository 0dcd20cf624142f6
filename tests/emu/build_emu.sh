#!/bin/bash
# TEST-ONLY: compiles the kernel sources with g++ against tests/emu/cuda_emu.h (see that header).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
src="$here/../../evolutionary_illusion_generator_b200/csrc"
mkdir -p "$here/_build"
g++ -O2 -g -std=c++17 -ffp-contract=off -fPIC -shared -x c++ -I"$here" -I"$src" -DEIG_EMU \
    -o "$here/_build/libeig_emu.so" "$src/eig_api.cu"
echo "built $here/_build/libeig_emu.so"
