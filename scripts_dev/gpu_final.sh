#!/bin/bash
# round-end style verification: GPU tests, smoke, default bench (both arms), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"; cut -c1-260 gpurun_out/bench_final_ref.json
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-gnn > gpurun_out/ncu_l6.log 2>&1; echo "ncu rc=$?"
