#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu17.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu17.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gnn --no-e2e > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"; tail -5 gpurun_out/bench17.err; python -c "
import json; d=json.load(open('gpurun_out/bench17.json')); print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'])"
timeout 100 python scripts_dev/slam_phases.py 2>&1 | tail -14
