mkdir -p gpurun_out
ls -la --time-style=full-iso drl_graph_exploration_b200/libdge.so | cut -c30-
timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|parity|steps=" gpurun_out/pytest_gpu.log | tail -8; grep -E "Error" gpurun_out/pytest_gpu.log | cut -c1-1200 | head -6
timeout 300 python scripts_dev/slam_phases.py 80 > gpurun_out/slam_phases.log 2>&1; cat gpurun_out/slam_phases.log | tail -14
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; cat gpurun_out/bench5.json | cut -c1-1800; tail -3 gpurun_out/bench5.err
