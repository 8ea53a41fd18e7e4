#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_loop_gpu.py -x -q > gpurun_out/pytest_host.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_host.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['ms_per_launch'])"
timeout 120 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-overlap > gpurun_out/bench2_noov.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench2_noov.json')); print('no-overlap', d['value'], d['ms_per_step'], d['e2e'])"
timeout 120 python scripts_dev/slam_phases.py 80 > gpurun_out/slam_phases.txt 2>&1; cat gpurun_out/slam_phases.txt
timeout 120 python scripts_dev/host_profile.py > gpurun_out/host_profile.txt 2>&1; cat gpurun_out/host_profile.txt
