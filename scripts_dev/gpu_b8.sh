#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py tests/test_engine_csr_gpu.py -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu8.log
timeout 300 python bench.py --workload gnn --steps 30 --warmup 5 > gpurun_out/bench_gnn.json 2> gpurun_out/bench_gnn.err; echo "gnn rc=$?"; tail -5 gpurun_out/bench_gnn.err; cat gpurun_out/bench_gnn.json
timeout 300 python bench.py --workload gnn --steps 30 --warmup 5 --train-gemm tc3 > gpurun_out/bench_gnn_tc3.json 2> gpurun_out/bench_gnn_tc3.err; echo "gnn rc=$?"; tail -5 gpurun_out/bench_gnn_tc3.err; cat gpurun_out/bench_gnn_tc3.json
timeout 600 python bench.py --workload train --steps 30 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train rc=$?"; tail -15 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json
