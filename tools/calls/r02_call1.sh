#!/usr/bin/env bash
# round-2 call 1: GPU tests with the round's first fixes, strong-scaling proxy on one GPU (4096/N replications),
# secondary configs, and ncu --set full of the N, U0 and Uniform-W kernels (unprofiled in round 1)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/c1_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc $?" >> $O/c1_pytest.log
for reps in 512 1024 2048 4096; do
  for shape in small big; do
    RACH_BLOCK=$shape python tools/ncu_target.py --reps $reps --runs 2 >> $O/c1_strong_proxy.txt 2>&1
  done
done
python tools/bench_configs.py > $O/c1_bench_configs.json 2> $O/c1_bench_configs.err
T="python tools/ncu_target.py --variant n --nue 50000 --reps 1024"
$T > $O/c1_plain_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel_n -c 1 -o $O/r02_prof_n $T > $O/c1_ncu_n.log 2>&1
T="python tools/ncu_target.py --variant u0 --nue 100000 --reps 1024"
$T > $O/c1_plain_u0.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_u0_kernel -c 1 -o $O/r02_prof_u0 $T > $O/c1_ncu_u0.log 2>&1
T="python tools/ncu_target.py --variant w --distribution 1 --nue 100000 --reps 256"
$T > $O/c1_plain_uni.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02_prof_uniform $T > $O/c1_ncu_uni.log 2>&1
python tools/phase_profile.py --distribution 1 --reps 256 > $O/c1_phase_uniform.json 2>&1
tail -3 $O/c1_pytest.log; cat $O/c1_strong_proxy.txt
