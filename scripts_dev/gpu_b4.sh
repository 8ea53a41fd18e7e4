#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"; tail -5 gpurun_out/bench4.err; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print(d['value'], d['ms_per_step'], d['gnn_graphs_per_s'], d['e2e'], d['roofline']['ms_per_launch'])"
timeout 120 python scripts_dev/host_loop_profile.py > gpurun_out/host_loop_profile.txt 2>&1; cat gpurun_out/host_loop_profile.txt
timeout 300 python scripts_dev/c4_sweep.py > gpurun_out/c4_sweep.md 2>&1; cat gpurun_out/c4_sweep.md
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32x3 -s 3 -c 1 -f -o gpurun_out/prof_gemm python scripts_dev/gemm_bench.py > gpurun_out/ncu_g.log 2>&1; tail -3 gpurun_out/ncu_g.log
