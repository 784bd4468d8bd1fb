#!/bin/bash
# A/B on one box: new = working tree, old = cilqr_b200/lib/libcilqr_b200_old.so (the previous commit, built by hand);
# then the parity / host-path suites on the new build
mkdir -p gpurun_out
cp cilqr_b200/lib/libcilqr_b200.so /tmp/new.so
for v in new old new old; do
  if [ $v = old ]; then cp cilqr_b200/lib/libcilqr_b200_old.so cilqr_b200/lib/libcilqr_b200.so; else cp /tmp/new.so cilqr_b200/lib/libcilqr_b200.so; fi
  python bench.py --steps 6 --warmup 3 --no-corridor --no-dp --no-latency --no-e2e --cpu-sample 1024 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['value']), round(d['value_one_in_flight']), round(d['roofline']['kernel_ms'],2), d['config']['parity']['worst'], d['config']['parity']['identical_path'])"
done | tee gpurun_out/r2n_ab.log
cp /tmp/new.so cilqr_b200/lib/libcilqr_b200.so
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostpath.py tests/test_gpu_strict.py -m gpu -x -q 2>&1 | tail -3
