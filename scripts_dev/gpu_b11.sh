#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edge_cases_gpu.py tests/test_graph_gnn_parity_gpu.py -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu11.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu11.log
timeout 200 python scripts_dev/ggnn_bench.py 2>&1 | tail -8
