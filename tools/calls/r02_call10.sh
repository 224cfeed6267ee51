#!/usr/bin/env bash
# round-2 call 10: medium block shapes for few replications per SM (strong scaling), N after the occasion-counter fix,
# wide parity fuzz over shapes / point views
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
echo "== 512 reps (8-GPU strong share), RACH_BLOCK=big, by big-shape variant"
RACH_BLOCK=big python tools/ncu_target.py --reps 512 --runs 2
for v in big288x4 big320x4 big384x3 big224x5 big192x6; do echo $v; RACH_BLOCK=big RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 512 --runs 2; done
echo "== 1024 reps (4-GPU strong share): small vs variants"
python tools/ncu_target.py --reps 1024 --runs 2
for v in big224x5 big192x6; do echo $v; RACH_BLOCK=big RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 1024 --runs 2; done
echo "== N 50k x 1024 / 2048"
python tools/ncu_target.py --variant n --nue 50000 --reps 1024 --runs 2
python tools/ncu_target.py --variant n --nue 50000 --reps 2048 --runs 2
} > $O/c10_timings.txt 2>&1
python tools/gpu_fuzz.py 150 20261 > $O/c10_fuzz.txt 2>&1
cat $O/c10_timings.txt; tail -3 $O/c10_fuzz.txt
