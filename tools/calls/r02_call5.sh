#!/usr/bin/env bash
# round-2 call 5: light-ms path (one warp, no block barrier): GPU suite, headline / Uniform / light Beta with and without it
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c5_pytest.log 2>&1; echo "pytest rc $?" >> $O/c5_pytest.log
{
echo "== headline 4096 reps: default (light path on) / nolight / b8"
python tools/ncu_target.py --reps 4096 --runs 2
for v in nolight b8; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
echo "== uniform 100k x 256: default / nolight"
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/nolight.so python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
echo "== beta 10k x 4096: default / nolight"
python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/nolight.so python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2
echo "== strong proxy"
for reps in 512 1024 2048; do python tools/ncu_target.py --reps $reps --runs 2; done
} > $O/c5_timings.txt 2>&1
python tools/bench_configs.py > $O/c5_bench_configs.json 2> $O/c5_bench_configs.err
T="python tools/ncu_target.py --distribution 1 --reps 256"
$T > $O/c5_plain_uni.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02d_prof_uniform_light $T > $O/c5_ncu_uni.log 2>&1
tail -4 $O/c5_pytest.log; cat $O/c5_timings.txt
