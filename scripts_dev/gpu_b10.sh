#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q -s -x --durations=8 > gpurun_out/pytest_edge.log 2>&1; echo "edge rc=$?"; tail -40 gpurun_out/pytest_edge.log
timeout 900 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_edge_cases_gpu.py > gpurun_out/pytest_gpu10.log 2>&1; echo "all rc=$?"; tail -16 gpurun_out/pytest_gpu10.log
