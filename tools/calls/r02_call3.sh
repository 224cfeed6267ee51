#!/usr/bin/env bash
# round-2 call 3: GPU tests (streamed dumps, binary logs, aggregation), headline with the 128 x 9 shape and mover
# micro-changes vs variants, bench.py both arms, traffic capture at the bench workload, launch list, secondary configs
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/c3_pytest.log 2>&1; echo "pytest rc $?" >> $O/c3_pytest.log
{
echo "== headline 4096 reps: default (128x9) and variants"
python tools/ncu_target.py --reps 4096 --runs 2
for v in b8 b10 ilp3b9; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/ncu_target.py --reps 4096 --runs 2; done
echo "== strong proxy, automatic shape"
for reps in 512 1024 2048; do python tools/ncu_target.py --reps $reps --runs 2; done
echo "== uniform 100k x 256: default / 64-thread blocks"
python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/t64b16.so python tools/ncu_target.py --distribution 1 --reps 256 --runs 2
echo "== N 50k"
python tools/ncu_target.py --variant n --nue 50000 --reps 1024 --runs 2
python tools/ncu_target.py --variant n --nue 50000 --reps 2048 --runs 2
} > $O/c3_timings.txt 2>&1
( time python bench.py > $O/c3_bench_default.json 2> $O/c3_bench_default.err ) 2> $O/c3_bench_time.txt
( time python bench.py --impl reference --steps 2 --warmup 1 > $O/c3_bench_reference.json 2> $O/c3_bench_reference.err ) 2>> $O/c3_bench_time.txt
python tools/bench_configs.py > $O/c3_bench_configs.json 2> $O/c3_bench_configs.err
T="python tools/ncu_target.py --reps 4096"
$T > $O/c3_plain_4096.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ra_step_kernel -c 1 -o $O/r02c_traffic_4096 $T > $O/c3_ncu_traffic.log 2>&1
T="python tools/ncu_target.py --reps 1332"
$T > $O/c3_plain_1332.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02c_prof_w_128x9 $T > $O/c3_ncu_w.log 2>&1
T="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$T > $O/c3_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02c_launches_bench.csv $T > $O/c3_ncu_bench.log 2>&1
tail -3 $O/c3_pytest.log; cat $O/c3_timings.txt; cat $O/c3_bench_time.txt
