ls -la --time-style=full-iso drl_graph_exploration_b200/libdge.so | cut -c30-
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -3 gpurun_out/bench6.err; cat gpurun_out/bench6.json
