#!/bin/bash
set -x
mkdir -p gpurun_out/j7
O=gpurun_out/j7
# one PredNet step (11 conv launches with folding) of C3, full sections + source counters
EIG_FOLD=1 EIG_NO_GRAPH=1 EIG_NO_OVERLAP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc -s 33 -c 11 \
   -o $O/tc_c3_fold_step python profiles/experiments/one_eval.py --workload c3 --evals 1 > $O/ncu_fold.log 2>&1
EIG_FOLD=1 EIG_PRECISION=2 EIG_NO_GRAPH=1 EIG_NO_OVERLAP=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc -s 33 -c 11 \
   -o $O/tc_c3_fold_step_p2 python profiles/experiments/one_eval.py --workload c3 --evals 1 > $O/ncu_fold_p2.log 2>&1
ls -la $O
