#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 scripts_dev/a2c_n2_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS" | tail -4
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --workload train --steps 40 --warmup 5 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train rc=$?"; tail -1 gpurun_out/bench_train_n2.json | cut -c1-200
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_policy_n2.json 2> gpurun_out/bench_policy_n2.err; echo "policy rc=$?"; tail -3 gpurun_out/bench_policy_n2.err | cut -c1-200; tail -1 gpurun_out/bench_policy_n2.json | cut -c1-200
timeout 200 $TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 2>/dev/null | cut -c1-200
