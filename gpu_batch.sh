ls -la --time-style=full-iso drl_graph_exploration_b200/libdge.so | cut -c30-
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 120 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -3 gpurun_out/bench8.err; cut -c1-230 gpurun_out/bench8.json; grep -o "\"ms_per_launch[^}]*}" gpurun_out/bench8.json
