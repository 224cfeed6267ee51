#!/usr/bin/env bash
# round-2 call 2: parity of every kernel instantiation, fixed-family view vs runtime view on the headline workload,
# block shapes at 4096/N replications, N with the warp-form sector decision, tuning variants, fresh ncu captures
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/c2_pytest.log 2>&1; echo "pytest rc $?" >> $O/c2_pytest.log
{
echo "== headline 4096 reps: fixed view / runtime view"
python tools/ncu_target.py --reps 4096 --runs 2
RACH_FIXED=0 python tools/ncu_target.py --reps 4096 --runs 2
echo "== strong proxy (fixed view): reps x shape"
for reps in 512 1024 2048; do for shape in small big huge; do echo "shape $shape"; RACH_BLOCK=$shape python tools/ncu_target.py --reps $reps --runs 2; done; done
echo "== automatic shape"
for reps in 256 512 1024 2048; do python tools/ncu_target.py --reps $reps --runs 2; done
echo "== tuning variants, 4096 reps"
for v in ilp3 ilp1 b9 b10 t96b10; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
echo "== N 50k x 1024 / x 2048"
python tools/ncu_target.py --variant n --nue 50000 --reps 1024 --runs 2
python tools/ab_n.py default 5g-nr-randomaccess_b200/tune/n128x8.so 5g-nr-randomaccess_b200/tune/n256x4.so 5g-nr-randomaccess_b200/tune/n96x10.so
echo "== uniform 100k x 256 (auto shape), forced shapes"
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
for shape in small big huge; do RACH_BLOCK=$shape python tools/ncu_target.py --distribution 1 --reps 256 --runs 2; done
} > $O/c2_timings.txt 2>&1
T="python tools/ncu_target.py --reps 1184"
$T > $O/c2_plain_w.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02b_prof_w_fixed $T > $O/c2_ncu_w.log 2>&1
T="python tools/ncu_target.py --variant n --nue 50000 --reps 1024"
$T > $O/c2_plain_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel_n -c 1 -o $O/r02b_prof_n $T > $O/c2_ncu_n.log 2>&1
tail -3 $O/c2_pytest.log; cat $O/c2_timings.txt
