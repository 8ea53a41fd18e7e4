#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"; tail -5 gpurun_out/bench6.err; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print(d['value'], d['ms_per_step'], d['gnn_graphs_per_s'], d['e2e']['value'], d['roofline']['ms_per_launch'])"
timeout 300 python scripts_dev/c4_sweep.py > gpurun_out/c4_sweep.md 2>&1; cat gpurun_out/c4_sweep.md
