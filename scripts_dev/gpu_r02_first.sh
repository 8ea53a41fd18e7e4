#!/bin/bash
# First GPU call of round 2: the parity suite as the driver runs it (with the printed row counts of the golden episodes, three of which
# have not been on a GPU yet), then the default bench line of both arms.
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts_dev/gpu_r02_first.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 700 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
grep -h "rows of the reference" gpurun_out/r02_pytest_gpu.log | cut -c1-160
timeout 400 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-200 gpurun_out/r02_bench_ref.json
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench.json
