#!/usr/bin/env bash
# round-2 call 11: N kernel with 4 + 1 instead of 5 + 2 block barriers per occasion / Msg3 ms: GPU suite + timing
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c11_pytest.log 2>&1; echo "pytest rc $?" >> $O/c11_pytest.log
{
echo "== N 50k x 1024 / 2048 / 100k x 1024"
python tools/ncu_target.py --variant n --nue 50000 --reps 1024 --runs 2
python tools/ncu_target.py --variant n --nue 50000 --reps 2048 --runs 2
python tools/ncu_target.py --variant n --nue 100000 --reps 1024 --runs 2
echo "== uniform phase profile"
python tools/phase_profile.py --distribution 1 --reps 256
} > $O/c11_timings.txt 2>&1
T="python tools/ncu_target.py --variant n --nue 50000 --reps 1024"
$T > $O/c11_plain_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel_n -c 1 -o $O/r02h_prof_n_final $T > $O/c11_ncu_n.log 2>&1
tail -3 $O/c11_pytest.log; cat $O/c11_timings.txt
