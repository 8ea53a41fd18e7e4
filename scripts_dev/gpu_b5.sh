#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; tail -5 gpurun_out/bench5.err; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print(d['value'], d['ms_per_step'], d['gnn_graphs_per_s'], d['e2e'], d['roofline']['ms_per_launch'])"
timeout 120 python scripts_dev/host_loop_profile.py > gpurun_out/host_loop_profile.txt 2>&1; cat gpurun_out/host_loop_profile.txt
timeout 300 python scripts_dev/c4_sweep.py > gpurun_out/c4_sweep.md 2>&1; cat gpurun_out/c4_sweep.md; timeout 100 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-flush-l2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(\"noflush\", d[\"value\"], d[\"ms_per_step\"], d[\"roofline\"][\"ms_per_launch\"])"
