#!/usr/bin/env bash
# round-2 call 21: A/B of the always-atomic cohort minimum, loop unrolling, block shapes with more L1 per SM
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
echo "== headline 4096 (2 runs each, last one printed)"
python tools/ncu_target.py --reps 4096 --runs 2
for v in aminalways unroll2 nt192b6 nt160b7 nt256b4 aminalways; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
python tools/ncu_target.py --reps 4096 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/aminalways.so python tools/ncu_target.py --reps 4096 --runs 2
} > $O/c21_timings.txt 2>&1
cat $O/c21_timings.txt
