#!/bin/bash
# First GPU call of round 2: (1) the parity suite as the driver runs it, (2) the staged experiment (bucketed ticks): opt-in bit-exact test,
# then the A/B bench lines (lock step vs bucketed, static and adaptive), (3) the default bench line of both arms.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts_dev/gpu_r02_first.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
grep -h "rows of the reference" gpurun_out/r02_pytest_gpu.log | cut -c1-160
DGE_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_bucketed_loop_gpu.py -x -q -s > gpurun_out/r02_pytest_bucketed.log 2>&1; echo "bucketed test rc=$?"; tail -15 gpurun_out/r02_pytest_bucketed.log | cut -c1-300
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gnn --no-e2e > gpurun_out/r02_ab_lockstep.json 2> gpurun_out/r02_ab_lockstep.err; cut -c1-200 gpurun_out/r02_ab_lockstep.json
for r in 2 3 4; do
  timeout 300 python bench.py --workload policy-bucketed --steps 100 --warmup 20 --short-ticks $r > gpurun_out/r02_ab_bucketed_r$r.json 2> gpurun_out/r02_ab_bucketed_r$r.err; echo "r=$r rc=$?"; cut -c1-200 gpurun_out/r02_ab_bucketed_r$r.json; tail -2 gpurun_out/r02_ab_bucketed_r$r.err | cut -c1-300
done
timeout 300 python bench.py --workload policy-bucketed --steps 100 --warmup 20 --adaptive-ticks --short-fraction 0.85 > gpurun_out/r02_ab_bucketed_adaptive.json 2> gpurun_out/r02_ab_bucketed_adaptive.err; cut -c1-300 gpurun_out/r02_ab_bucketed_adaptive.json
timeout 400 python bench.py --impl reference --steps 200 --warmup 20 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-200 gpurun_out/r02_bench_ref.json
timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02_bench.json
