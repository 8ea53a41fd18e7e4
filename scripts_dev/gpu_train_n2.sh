#!/bin/bash
# DQN training (C3) at N = 2: native path overlapped / sequential, autograd path (A/B)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for v in "" "--no-overlap" "--train-gemm fp32"; do
  timeout 300 $TR bench.py --gpus 2 --workload train --steps 60 --warmup 10 $v 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('N=2 [$v]', round(d['value']), 'env-steps/s', round(d['ms_per_step'], 2), 'ms/tick', 'train steps/s/rank', round(d['train_steps_per_s'], 1), 'allreduce us', d.get('allreduce_us'))
"
done
timeout 300 python bench.py --workload train --steps 60 --warmup 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('N=1', round(d['value']), 'env-steps/s', round(d['ms_per_step'], 2), 'ms/tick')
"
