#!/usr/bin/env bash
# round-2 call 23: A/B of branch removals in the mover tail
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
echo "== headline 4096 (2 runs each, last one printed); order: default $*"
python tools/ncu_target.py --reps 4096 --runs 2
for v in "$@"; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
python tools/ncu_target.py --reps 4096 --runs 2
} > $O/c23_timings.txt 2>&1
cat $O/c23_timings.txt
