#!/bin/bash
# round-end style verification (session 3): GPU tests, smoke, default bench (both arms), ncu launch lists of the bench command and of one GG-NN training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final_s3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final_s3.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/bench_final_s3_ref.json 2> gpurun_out/bench_final_s3_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/bench_final_s3_ref.json
timeout 600 python bench.py > gpurun_out/bench_final_s3.json 2> gpurun_out/bench_final_s3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final_s3.err; cut -c1-400 gpurun_out/bench_final_s3.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ggnn_train_s3b.csv python scripts_dev/ggnn_train_step_once.py > gpurun_out/ncu_ggnn_s3b.log 2>&1; echo "ncu ggnn rc=$?"
