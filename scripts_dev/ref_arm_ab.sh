#!/bin/bash
# reference arm: effect of the OpenMP environment / torch thread count on the CPU-oracle + torch-CPU GCN loop (no GPU work)
run() { python - "$@" <<'PY'
import os, sys, time
sys.path.insert(0, "/root/repo")
gt = int(sys.argv[1])
import torch
from oracle.cpu_loop import cpu_reference
import oracle.cpu_loop as cl
run_tick, count = cpu_reference(20, 30, 256, os.cpu_count(), 192, gnn_threads=gt)
for _ in range(150): run_tick()
c0, t0 = count(), time.perf_counter()
n = 0
while time.perf_counter() - t0 < 8: run_tick(); n += 1
print(f"OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')} gnn_threads={gt} torch threads={torch.get_num_threads()}: {(count() - c0) / (time.perf_counter() - t0):.0f} env-steps/s, {1e3 * (time.perf_counter() - t0) / n:.1f} ms/tick")
PY
}
OMP_NUM_THREADS=16 MKL_NUM_THREADS=16 run 16
run 16
run 8
run 4
OMP_NUM_THREADS=1 run 16
