#!/usr/bin/env bash
# round-2 call 19: re-transmitters parked in registers (RaPend), closed-form slot alignment, 32-bit bucket index:
# GPU suite on the new default, A/B timings against the committed kernel, short wide fuzz
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c19_pytest.log 2>&1; echo "pytest rc $?" >> $O/c19_pytest.log
{
echo "== headline 4096 (2 runs each, last one printed): default = defer"
python tools/ncu_target.py --reps 4096 --runs 2
for v in base nodefer minb10 minb10nd base; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
python tools/ncu_target.py --reps 4096 --runs 2
echo "== 50k x 2048: default / base"
python tools/ncu_target.py --nue 50000 --reps 2048 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/base.so python tools/ncu_target.py --nue 50000 --reps 2048 --runs 2
echo "== strong-scaling share 512 x 100k: default / base"
python tools/ncu_target.py --reps 512 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/base.so python tools/ncu_target.py --reps 512 --runs 2
echo "== uniform 100k x 256: default / base"
python tools/ncu_target.py --reps 256 --distribution 1 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/base.so python tools/ncu_target.py --reps 256 --distribution 1 --runs 2
} > $O/c19_timings.txt 2>&1
python tools/gpu_fuzz.py 150 9001 > $O/c19_fuzz.txt 2>&1
tail -3 $O/c19_pytest.log; cat $O/c19_timings.txt; tail -3 $O/c19_fuzz.txt
