#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_api_gpu.py -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu12.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu12.log
for k in k_vmap_env k_slam k_aggregate k_gcn_conv_small k_graph_fill; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -f -o gpurun_out/prof_bench_$k python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
