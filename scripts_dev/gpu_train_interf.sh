#!/bin/bash
# which earlier leg of the default bench line slows the train_c3 leg down (same process)
for v in "--no-e2e --no-gnn --no-c4 --no-cpu-baseline" "--no-e2e --no-c4 --no-cpu-baseline"; do
  timeout 400 python bench.py --steps 4 --warmup 3 --preroll 100 $v 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('[$v] train_c3 ms/tick', round(d['train_c3']['ms_per_step'], 2), 'host issue', round(d['train_c3']['host_issue_ms_per_tick'], 2), d['train_c3']['allocator'])
"
done
