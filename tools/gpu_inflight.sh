#!/bin/bash
# value with 2, 3 and 4 batches in flight (device-resident legs only)
mkdir -p gpurun_out
for f in 2 3 4 2 3; do
  python bench.py --steps 6 --warmup 3 --in-flight $f --no-corridor --no-dp --no-latency --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('in_flight $f', round(d['value']), round(d['value_one_in_flight']), round(d['ms_per_step'],2))"
done | tee gpurun_out/inflight.log
