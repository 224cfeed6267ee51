#!/usr/bin/env bash
# round-2 call 22: full-set ncu capture of the step kernel after the mover-path changes (one wave, 1332 replications)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T="python tools/ncu_target.py --reps 1332"
$T > $O/c22_plain_1332.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02k_prof_w $T > $O/c22_ncu_w.log 2>&1
cat $O/c22_plain_1332.log; tail -3 $O/c22_ncu_w.log
