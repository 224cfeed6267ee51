#!/usr/bin/env bash
# round-2 call 32: which change slowed the runtime point view (configs[4] grid: 1875 -> 2013 ms)?
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
{
python tools/grid_target.py 128
for v in "$@"; do echo $v; RACH_GPU_LIB=5g-nr-randomaccess_b200/tune/$v.so python tools/grid_target.py 128; done
python tools/grid_target.py 128
} > $O/c32_grid.txt 2>&1
cat $O/c32_grid.txt
