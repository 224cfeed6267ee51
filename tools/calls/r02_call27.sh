#!/usr/bin/env bash
# round-2 call 27: A/B of the phase 0 / phase 5 changes at medium load
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
for args in "--reps 1332 --distribution 1" "--reps 4096 --nue 10000" "--reps 2048 --nue 20000" "--reps 2048 --nue 30000" "--reps 256 --distribution 1"; do
  echo "== $args: default / f5cc2ae / norowskip / default"
  python tools/ncu_target.py $args --runs 3
  RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/f5cc2ae.so python tools/ncu_target.py $args --runs 3
  RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/norowskip.so python tools/ncu_target.py $args --runs 3
  python tools/ncu_target.py $args --runs 3
done
} > $O/c27_timings.txt 2>&1
cat $O/c27_timings.txt
