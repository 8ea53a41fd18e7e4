#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu20.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu20.log
timeout 300 python bench.py --workload gnn --steps 30 --warmup 5 > gpurun_out/bench_gnn20.json 2> gpurun_out/bench_gnn20.err; echo "gnn rc=$?"; tail -3 gpurun_out/bench_gnn20.err; python -c "
import json; d=json.load(open('gpurun_out/bench_gnn20.json')); print(d['forward'], d['forward_backward'])"
