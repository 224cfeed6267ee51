#!/usr/bin/env bash
# round-2 8-GPU call 2: BASELINE configs[2] / configs[4] at full size through ra_sim_create(devices[0..7]); the README
# sweep from the reference-compatible CLI over 8 devices with binary per-UE logs for one point
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python tools/bench_configs_multi.py 8 > $O/c16_configs_8gpu.json 2> $O/c16_configs_8gpu.err
( time 5g-nr-randomaccess_b200/host/rach_sim -t 100 --no-logs --devices 0,1,2,3,4,5,6,7 --outdir /tmp/sweep8 ) > $O/c16_cli_sweep.txt 2>&1
python 5g-nr-randomaccess_b200/average_performance.py --from-files /tmp/sweep8/NomaBetaResults --out $O/c16_results_from_cli_files.csv > $O/c16_results_from_cli_files.log 2>&1
( time 5g-nr-randomaccess_b200/host/rach_sim -t 64 --nue 100000 --binlog --devices 0,1,2,3,4,5,6,7 --outdir /tmp/logs8 ) > $O/c16_cli_binlog.txt 2>&1
ls -la /tmp/logs8/NomaBetaResults | head -5 >> $O/c16_cli_binlog.txt; du -sh /tmp/logs8 >> $O/c16_cli_binlog.txt
python -m pytest tests -m gpu -q -k "multi_device or binary_logs or device_list" > $O/c16_pytest.log 2>&1; echo "pytest rc $?" >> $O/c16_pytest.log
cat $O/c16_configs_8gpu.json | head -30; grep -E "real|rach_sim:" $O/c16_cli_sweep.txt $O/c16_cli_binlog.txt; tail -3 $O/c16_pytest.log; tail -14 $O/c16_results_from_cli_files.log
