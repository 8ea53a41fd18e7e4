#!/bin/bash
# full parity suite, fresh ncu capture of k_slam (roofline.traffic), launch list of the native training step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_slam -s 650 -c 1 -f -o gpurun_out/prof_slam_r02c python bench.py --steps 2 --warmup 3 --preroll 600 --no-e2e --no-cpu-baseline --no-gnn --no-c4 --no-train > gpurun_out/ncu_s.log 2>&1; tail -1 gpurun_out/ncu_s.log | cut -c1-100
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 200 --csv --log-file gpurun_out/r02_launches_train.csv python bench.py --workload gnn --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_t.log 2>&1; echo "ncu train rc=$?"
