#!/usr/bin/env bash
# round-2 call 24: full-set ncu capture of the step kernel on the Uniform 100k x 256 workload (configs[1])
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T="python tools/ncu_target.py --reps 256 --distribution 1"
$T > $O/c28_plain_uni.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02m_prof_uniform $T > $O/c28_ncu_uni.log 2>&1
cat $O/c28_plain_uni.log; tail -3 $O/c28_ncu_uni.log
