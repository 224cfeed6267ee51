#!/usr/bin/env bash
# round-2 call 20: which of the mover-path micro changes pay (A/B), and one bounded try of compute-sanitizer
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
echo "== headline 4096 (2 runs each, last one printed)"
python tools/ncu_target.py --reps 4096 --runs 2
for v in noalign noidx32 aminalways nodefer base; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
python tools/ncu_target.py --reps 4096 --runs 2
} > $O/c20_timings.txt 2>&1
{
echo "== compute-sanitizer memcheck (bounded)"
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py; echo "memcheck rc $?"
echo "== compute-sanitizer racecheck (bounded)"
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py; echo "racecheck rc $?"
} > $O/c20_sanitizer.txt 2>&1
cat $O/c20_timings.txt; tail -40 $O/c20_sanitizer.txt
