#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; tail -5 gpurun_out/bench7.err; cat gpurun_out/bench7.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l4.log 2>&1; echo "ncu rc=$?"
