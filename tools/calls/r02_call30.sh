#!/usr/bin/env bash
# round-2 call 30: phase 5 and its barrier skipped when every singleton scan of the ms is answered: parity tests, A/B
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/c30_pytest.log 2>&1; echo "pytest rc $?" >> $O/c30_pytest.log
B=5g-nr-randomaccess_b200/tune/c5aa6b9d.so
{
for args in "--reps 256 --distribution 1" "--reps 1332 --distribution 1" "--reps 4096 --nue 10000" "--reps 2048 --nue 20000" "--reps 2048 --nue 30000" "--reps 4096"; do
  echo "== $args: default / previous commit / default"
  python tools/ncu_target.py $args --runs 3
  RACH_GPU_LIB=$B python tools/ncu_target.py $args --runs 3
  python tools/ncu_target.py $args --runs 3
done
} > $O/c30_timings.txt 2>&1
tail -3 $O/c30_pytest.log; cat $O/c30_timings.txt
