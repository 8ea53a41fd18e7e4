ls -la --time-style=full-iso drl_graph_exploration_b200/libdge.so | cut -c30-
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -3 gpurun_out/bench7.err; cut -c1-330 gpurun_out/bench7.json; grep -o '"gpu_launches.*' gpurun_out/bench7.json | cut -c1-700
timeout 600 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu-baseline --no-overlap > gpurun_out/bench7b.json 2> gpurun_out/bench7b.err; tail -3 gpurun_out/bench7b.err; cut -c1-330 gpurun_out/bench7b.json;  grep -o '"gpu_launches.*' gpurun_out/bench7b.json | cut -c1-700
