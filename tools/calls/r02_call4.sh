#!/usr/bin/env bash
# round-2 call 4: the full GPU suite (call 3 stopped at a test typo), block-shape variants for the headline and for the
# lightly loaded Uniform config, README sweep timing
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q > $O/c4_pytest.log 2>&1; echo "pytest rc $?" >> $O/c4_pytest.log
{
echo "== headline 4096 reps: default (128x9) and variants"
python tools/ncu_target.py --reps 4096 --runs 2
for v in b8 b10; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
echo "== uniform 100k x 256: default / 64-thread / 32-thread blocks"
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
for v in t64b16 t32b16; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --distribution 1 --reps 256 --runs 2; done
echo "== beta 10k x 4096 (light, BASELINE configs[0] size): default / 64 / 32"
python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2
for v in t64b16 t32b16; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2; done
echo "== retx 50, 100k x 1332"
python tools/ncu_target.py --reps 1332 --retx 50 --runs 2
echo "== U0 100k x 1024"
python tools/ncu_target.py --variant u0 --reps 1024 --runs 2
} > $O/c4_timings.txt 2>&1
python tools/bench_configs.py > $O/c4_bench_configs.json 2> $O/c4_bench_configs.err
tail -4 $O/c4_pytest.log; cat $O/c4_timings.txt
