#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu9.log
timeout 120 python scripts_dev/host_loop_profile.py > gpurun_out/host_loop_profile_v2.txt 2>&1; cat gpurun_out/host_loop_profile_v2.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench rc=$?"; tail -5 gpurun_out/bench9.err; python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print(d['value'], d['ms_per_step'], d['gnn_graphs_per_s'], d['e2e'], d['roofline']['ms_per_launch'])"
