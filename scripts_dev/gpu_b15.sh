#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_host_loop_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu15.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu15.log
timeout 120 python scripts_dev/host_loop_profile.py > gpurun_out/host_loop_profile_v3.txt 2>&1; cat gpurun_out/host_loop_profile_v3.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gnn > gpurun_out/bench15.json 2> gpurun_out/bench15.err; echo "bench rc=$?"; tail -5 gpurun_out/bench15.err; python -c "
import json; d=json.load(open('gpurun_out/bench15.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
