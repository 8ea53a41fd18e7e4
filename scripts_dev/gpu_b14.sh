#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"; tail -5 gpurun_out/bench14.err; cat gpurun_out/bench14.json
for k in k_aggregate k_gemm_tf32x3 k_gcn_conv_small; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o gpurun_out/prof_c5_$k python bench.py --workload gnn --steps 3 --warmup 3 > gpurun_out/ncu_c5_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload gnn --steps 4 --warmup 3 > gpurun_out/ncu_lc5.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/prof_c5*
