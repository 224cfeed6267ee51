#!/usr/bin/env bash
# round-2 call 8: GPU suite after the device-list fix, ncu of the Uniform config with light path v2, README tables
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/c8_pytest.log 2>&1; echo "pytest rc $?" >> $O/c8_pytest.log
python 5g-nr-randomaccess_b200/average_performance.py --readme-tables --seeds 100 > $O/c8_readme_tables.md 2> $O/c8_readme_tables.err
python 5g-nr-randomaccess_b200/average_performance.py --seeds 100 --out $O/c8_results.csv > $O/c8_results.log 2>&1
T="python tools/ncu_target.py --distribution 1 --reps 256"
$T > $O/c8_plain_uni.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ra_step_kernel -c 1 -o $O/r02f_prof_uniform_light2 $T > $O/c8_ncu_uni.log 2>&1
tail -4 $O/c8_pytest.log
