#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --workload train --steps 30 --warmup 5 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train rc=$?"; tail -3 gpurun_out/bench_train_n2.err; tail -1 gpurun_out/bench_train_n2.json
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --workload gnn --steps 30 --warmup 5 > gpurun_out/bench_gnn_n2.json 2> gpurun_out/bench_gnn_n2.err; echo "gnn rc=$?"; tail -3 gpurun_out/bench_gnn_n2.err; tail -1 gpurun_out/bench_gnn_n2.json
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_policy_n2.json 2> gpurun_out/bench_policy_n2.err; echo "policy rc=$?"; tail -3 gpurun_out/bench_policy_n2.err; tail -1 gpurun_out/bench_policy_n2.json | cut -c1-600
