#!/bin/bash
set -x
mkdir -p gpurun_out/j16
O=gpurun_out/j16
EIG_NO_GRAPH=1 EIG_NO_OVERLAP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc -s 40 -c 2 \
   -o $O/tc_c3_l1 python profiles/experiments/one_eval.py --workload c3 --evals 1 > $O/ncu.log 2>&1
ls -la $O
