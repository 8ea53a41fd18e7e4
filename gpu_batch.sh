ls -la --time-style=full-iso drl_graph_exploration_b200/libdge.so | cut -c30-
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts_dev/slam_phases.py 80 > gpurun_out/slam_phases.log 2>&1; cat gpurun_out/slam_phases.log | tail -18 | grep -v "B0 sub"
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -3 gpurun_out/bench8.err; cut -c1-330 gpurun_out/bench8.json; grep -o "\"ms_per_launch[^}]*}" gpurun_out/bench8.json
