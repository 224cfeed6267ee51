#!/usr/bin/env bash
# round-2 call 26: phase 5 rank-by-counting for <= 32 singletons, phase 0 skips the empty cohorts of the window:
# GPU suite, then A/B against the previous commit
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c26_pytest.log 2>&1; echo "pytest rc $?" >> $O/c26_pytest.log
B=5g-nr-randomaccess_b200/tune/f5cc2ae.so
{
for args in "--reps 256 --distribution 1" "--reps 1332 --distribution 1" "--reps 4096 --nue 10000" "--reps 2048 --nue 20000" "--reps 4096"; do
  echo "== $args: default / previous commit / default"
  python tools/ncu_target.py $args --runs 3
  RACH_GPU_LIB=$B python tools/ncu_target.py $args --runs 3
  python tools/ncu_target.py $args --runs 3
done
} > $O/c26_timings.txt 2>&1
python tools/gpu_fuzz.py 100 777 > $O/c26_fuzz.txt 2>&1
tail -3 $O/c26_pytest.log; cat $O/c26_timings.txt; tail -2 $O/c26_fuzz.txt
