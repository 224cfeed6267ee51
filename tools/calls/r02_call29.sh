#!/usr/bin/env bash
# round-2 call 29: full-set ncu capture at medium load (Beta, 20k UEs x 1332 replications: the README sweep's regime)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T="python tools/ncu_target.py --reps 1332 --nue 20000"
$T > $O/c29_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02m_prof_beta20k $T > $O/c29_ncu.log 2>&1
cat $O/c29_plain.log; tail -2 $O/c29_ncu.log
