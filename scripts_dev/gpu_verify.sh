#!/bin/bash
# GPU verification batch: parity tests, bench, ncu launch list + full captures of the two dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 30 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slam -s 30 -c 1 -f -o gpurun_out/prof_slam python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_s.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vmap_cells -s 30 -c 1 -f -o gpurun_out/prof_vmap python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_v.log 2>&1
ls -la gpurun_out
