#!/usr/bin/env bash
# round-2 call 31: final captures of the round -- GPU suite, bench both arms, secondary configs, traffic + full-set ncu of the
# final W kernel, launch list of the bench command, wide fuzz
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c31_pytest.log 2>&1; echo "pytest rc $?" >> $O/c31_pytest.log
( time python bench.py --steps 20 --warmup 5 > $O/c31_bench_default.json 2> $O/c31_bench_default.err ) 2> $O/c31_bench_time.txt
( time python bench.py --impl reference --steps 3 --warmup 1 > $O/c31_bench_reference.json 2> $O/c31_bench_reference.err ) 2>> $O/c31_bench_time.txt
python tools/bench_configs.py > $O/c31_bench_configs.json 2> $O/c31_bench_configs.err
T="python tools/ncu_target.py --reps 4096"
$T > $O/c31_plain_4096.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ra_step_kernel -c 1 -o $O/r02n_traffic_4096 $T > $O/c31_ncu_traffic.log 2>&1
T="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$T > $O/c31_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02n_launches_bench.csv $T > $O/c31_ncu_bench.log 2>&1
T="python tools/ncu_target.py --reps 1332"
$T > $O/c31_plain_1332.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02n_prof_w_final $T > $O/c31_ncu_w.log 2>&1
python tools/gpu_fuzz.py 150 20261018 > $O/c31_fuzz.txt 2>&1
tail -3 $O/c31_pytest.log; cat $O/c31_bench_time.txt; cat $O/c31_plain_4096.log $O/c31_plain_1332.log; tail -2 $O/c31_fuzz.txt
