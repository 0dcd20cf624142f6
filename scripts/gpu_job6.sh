#!/bin/bash
set -x
mkdir -p gpurun_out/j6
O=gpurun_out/j6
for P in 7 4; do for F in 3 1 2; do
  EIG_TC_DBGFLAGS=$F timeout 300 tests/gpu/tc_check time c3 $P 2>&1 | cut -c1-330 > $O/tc_time_c3_p${P}_f$F.log
done; done
for F in 0 3 1 2; do
  EIG_TC_DBGFLAGS=$F timeout 300 tests/gpu/tc_check time 2>&1 | cut -c1-330 > $O/tc_time_c2_p7_f$F.log
done
EIG_TC_PASSES=4 EIG_TC_DBGFLAGS=0 timeout 300 tests/gpu/tc_check time 2>&1 | cut -c1-330 > $O/tc_time_c2_p4_f0.log
ls -la $O
