#!/bin/bash
# e2e (host buffers through cilqr_plan_batch) with 2, 3 and 4 host threads / handles in flight
mkdir -p gpurun_out
for f in 2 3 4 2 3; do
  python bench.py --steps 6 --warmup 3 --in-flight $f --no-corridor --no-dp --no-latency --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('in_flight $f value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('steps'))"
done | tee gpurun_out/e2e_inflight.log
