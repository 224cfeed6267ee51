#!/usr/bin/env bash
# round-2 call 17: streaming cache hints on the bucket records (A/B), long wide fuzz
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
echo "== headline 4096: default / stream hints (twice each, interleaved)"
python tools/ncu_target.py --reps 4096 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/stream.so python tools/ncu_target.py --reps 4096 --runs 2
python tools/ncu_target.py --reps 4096 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/stream.so python tools/ncu_target.py --reps 4096 --runs 2
echo "== 50k x 2048, grid config point (P64 BI40) 100k x 1332: default / stream"
python tools/ncu_target.py --nue 50000 --reps 2048 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/stream.so python tools/ncu_target.py --nue 50000 --reps 2048 --runs 2
} > $O/c17_timings.txt 2>&1
python tools/gpu_fuzz.py 420 4711 > $O/c17_fuzz.txt 2>&1
cat $O/c17_timings.txt; tail -3 $O/c17_fuzz.txt
