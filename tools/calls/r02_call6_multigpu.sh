#!/usr/bin/env bash
# round-2 multi-GPU call (gpurun --gpus 8): in-process multi-device tests, host CLI over all devices, weak and strong
# scaling of the bench workload at 2 / 4 / 8 GPUs (N = 1 for reference on the same box)
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/c6_gpus.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_host_cli.py -m gpu -q -k "multi_device or binary_logs or streamed" > $O/c6_pytest.log 2>&1; echo "pytest rc $?" >> $O/c6_pytest.log
( time 5g-nr-randomaccess_b200/host/rach_sim -t 128 --nue 100000 --no-logs --devices 0,1,2,3,4,5,6,7 --outdir /tmp/cli8 ) > $O/c6_cli8.txt 2>&1
( time 5g-nr-randomaccess_b200/host/rach_sim -t 128 --nue 100000 --no-logs --devices 0 --outdir /tmp/cli1 ) > $O/c6_cli1.txt 2>&1
cmp /tmp/cli8/NomaBetaResults/127_54_100000_Results.txt /tmp/cli1/NomaBetaResults/127_54_100000_Results.txt && echo "cli 8-device == 1-device result files" >> $O/c6_cli8.txt
python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/c6_weak_1.json 2> $O/c6_weak_1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 > $O/c6_weak_$n.json 2> $O/c6_weak_$n.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 3 --warmup 3 --scaling strong --reps 4096 > $O/c6_strong_$n.json 2> $O/c6_strong_$n.err
done
for f in $O/c6_weak_*.json $O/c6_strong_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['n_gpus'], d['scaling'], 'value %.4e' % d['value'], 'kernel_ms', round(d['kernel_ms_per_step'],1), 'ms_per_step', round(d['ms_per_step'],1))"; done
tail -3 $O/c6_pytest.log; tail -4 $O/c6_cli8.txt $O/c6_cli1.txt
