#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_env_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_golden_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu19.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu19.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gnn --no-e2e > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo "bench rc=$?"; tail -5 gpurun_out/bench19.err; python -c "
import json; d=json.load(open('gpurun_out/bench19.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
timeout 100 python scripts_dev/slam_phases.py 2>&1 | tail -16
