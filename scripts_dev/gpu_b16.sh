#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu16.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu16.log
timeout 400 python bench.py --workload train --steps 40 --warmup 5 > gpurun_out/bench_train_ovl.json 2> gpurun_out/bench_train_ovl.err; echo "train rc=$?"; tail -4 gpurun_out/bench_train_ovl.err; cut -c1-330 gpurun_out/bench_train_ovl.json; python -c "
import json; d=json.load(open('gpurun_out/bench_train_ovl.json')); print({k: d[k] for k in ('value','ms_per_step','decisions_per_s','train_steps_per_s','rollout')})"
timeout 400 python bench.py --workload train --steps 40 --warmup 5 --no-overlap > gpurun_out/bench_train_seq.json 2> gpurun_out/bench_train_seq.err; echo "train rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_train_seq.json')); print({k: d[k] for k in ('value','ms_per_step','decisions_per_s','train_steps_per_s','rollout')})"
