#!/bin/bash
# A/B on one box: rollout parking from the staged ring (new) vs from global memory (old)
mkdir -p gpurun_out
cp cilqr_b200/lib/libcilqr_b200.so /tmp/new.so
for v in new old new old; do
  if [ $v = old ]; then cp cilqr_b200/lib/libcilqr_b200_old.so cilqr_b200/lib/libcilqr_b200.so; else cp /tmp/new.so cilqr_b200/lib/libcilqr_b200.so; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-corridor --no-dp --no-latency --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['value']), round(d['value_one_in_flight']), round(d['roofline']['kernel_ms'],2))"
done | tee gpurun_out/r2n_ab.log
