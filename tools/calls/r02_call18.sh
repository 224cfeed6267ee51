#!/usr/bin/env bash
# round-2 call 18: cache-hint variants for the calendar records, new modulo: GPU suite, A/B timings
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c18_pytest.log 2>&1; echo "pytest rc $?" >> $O/c18_pytest.log
{
echo "== headline 4096: default (ld.cs + st.cs) and variants"
python tools/ncu_target.py --reps 4096 --runs 2
for v in nohints ldlu stcg stwt ldonly stonly; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
python tools/ncu_target.py --reps 4096 --runs 2
echo "== runtime point view (grid point P64 BI40) is not reachable from ncu_target; secondary configs follow"
} > $O/c18_timings.txt 2>&1
python tools/bench_configs.py > $O/c18_bench_configs.json 2> $O/c18_bench_configs.err
tail -3 $O/c18_pytest.log; cat $O/c18_timings.txt
