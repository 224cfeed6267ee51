#!/usr/bin/env bash
# round-2 call 12: end-of-run decision owned by warp 0 (race fix), N decision latch: GPU suite, fuzz, timings
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c12_pytest.log 2>&1; echo "pytest rc $?" >> $O/c12_pytest.log
python tools/gpu_fuzz.py 120 777 > $O/c12_fuzz.txt 2>&1
{
python tools/ncu_target.py --reps 4096 --runs 2
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
python tools/ncu_target.py --distribution 1 --reps 4096 --nue 20000 --runs 2
python tools/ncu_target.py --variant n --nue 50000 --reps 1024 --runs 2
python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2
} > $O/c12_timings.txt 2>&1
tail -3 $O/c12_pytest.log; tail -2 $O/c12_fuzz.txt; cat $O/c12_timings.txt
