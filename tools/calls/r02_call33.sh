#!/usr/bin/env bash
# round-2 call 33: parity suite + short fuzz on the final library (deferral only in the compile-time point view)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/c33_pytest.log 2>&1; echo "pytest rc $?" >> $O/c33_pytest.log
python tools/gpu_fuzz.py 50 33 > $O/c33_fuzz.txt 2>&1
tail -3 $O/c33_pytest.log; tail -2 $O/c33_fuzz.txt
