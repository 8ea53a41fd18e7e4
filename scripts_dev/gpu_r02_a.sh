#!/bin/bash
# Round-2 state check: parity suite, both bench arms with the driver's flags, launch list of the same command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-300 gpurun_out/r02_bench_ref.json
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench.err; cut -c1-3000 gpurun_out/r02_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --preroll 300 --no-e2e --no-cpu-baseline --no-gnn > gpurun_out/ncu_l.log 2>&1; echo "ncu rc=$?"
