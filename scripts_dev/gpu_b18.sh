#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vmap_kernel_gpu.py tests/test_env_parity_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu18.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu18.log
timeout 300 python scripts_dev/c4_sweep.py > gpurun_out/c4_sweep_v5.md 2>&1; cat gpurun_out/c4_sweep_v5.md
