#!/usr/bin/env bash
# round-2 call 7: light-ms path v2 (control block in registers, live-slot skip): GPU suite + timings (no profiler)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c7_pytest.log 2>&1; echo "pytest rc $?" >> $O/c7_pytest.log
{
echo "== headline 4096 reps: default / nolight"
python tools/ncu_target.py --reps 4096 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/nolight.so python tools/ncu_target.py --reps 4096 --runs 2
echo "== uniform 100k x 256: default / nolight"
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/nolight.so python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
echo "== beta 10k x 4096: default"
python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2
} > $O/c7_timings.txt 2>&1
python tools/bench_configs.py > $O/c7_bench_configs.json 2> $O/c7_bench_configs.err
tail -4 $O/c7_pytest.log; cat $O/c7_timings.txt
