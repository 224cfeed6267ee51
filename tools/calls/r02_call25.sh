#!/usr/bin/env bash
# round-2 call 25: light ms with one liveness ballot, phase timers as their own instantiation: A/B against the previous commit
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
B=5g-nr-randomaccess_b200/tune/c82e036.so
{
for args in "--reps 256 --distribution 1" "--reps 1332 --distribution 1" "--reps 4096 --nue 10000" "--reps 2048 --nue 20000" "--reps 4096"; do
  echo "== $args: default / previous commit / default"
  python tools/ncu_target.py $args --runs 3
  RACH_GPU_LIB=$B python tools/ncu_target.py $args --runs 3
  python tools/ncu_target.py $args --runs 3
done
} > $O/c25_timings.txt 2>&1
cat $O/c25_timings.txt
