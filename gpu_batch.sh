mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 4 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bench.log
