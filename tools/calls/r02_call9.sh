#!/usr/bin/env bash
# round-2 call 9: small-ms path (general phases by warp 0 alone): GPU suite, smoke, timings, bench both arms
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c9_pytest.log 2>&1; echo "pytest rc $?" >> $O/c9_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/c9_smoke.log 2>&1; echo "smoke rc $?" >> $O/c9_smoke.log
{
echo "== headline 4096 reps: default / nolight"
python tools/ncu_target.py --reps 4096 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/nolight.so python tools/ncu_target.py --reps 4096 --runs 2
echo "== uniform 100k x 256"
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
echo "== uniform 100k x 1332 (one wave)"
python tools/ncu_target.py --distribution 1 --reps 1332 --runs 2
echo "== beta 10k x 4096"
python tools/ncu_target.py --nue 10000 --reps 4096 --runs 2
echo "== strong proxy"
for reps in 512 1024 2048; do python tools/ncu_target.py --reps $reps --runs 2; done
} > $O/c9_timings.txt 2>&1
python tools/bench_configs.py > $O/c9_bench_configs.json 2> $O/c9_bench_configs.err
( time python bench.py --steps 20 --warmup 5 > $O/c9_bench_default.json 2> $O/c9_bench_default.err ) 2> $O/c9_bench_time.txt
( time python bench.py --impl reference --steps 3 --warmup 1 > $O/c9_bench_reference.json 2> $O/c9_bench_reference.err ) 2>> $O/c9_bench_time.txt
tail -3 $O/c9_pytest.log; cat $O/c9_smoke.log $O/c9_timings.txt $O/c9_bench_time.txt
