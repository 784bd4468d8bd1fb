#!/bin/bash
# A/B/C... on one box: every cilqr_b200/lib/variants/*.so is swapped in for libcilqr_b200.so in turn (two rounds), the
# bench's device-resident legs are timed, and the result fingerprint (parity.worst, identical paths) is printed.
# The shipped library is restored at the end.  usage: bash tools/gpu_ab.sh [rounds]
mkdir -p gpurun_out
cp cilqr_b200/lib/libcilqr_b200.so /tmp/shipped.so
for r in $(seq ${1:-2}); do
  for f in cilqr_b200/lib/variants/*.so; do
    v=$(basename $f .so)
    cp $f cilqr_b200/lib/libcilqr_b200.so
    python bench.py --steps 6 --warmup 3 --no-corridor --no-dp --no-latency --no-e2e --cpu-sample 1024 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['value']), round(d['value_one_in_flight']), round(d['roofline']['kernel_ms'],2), d['config']['parity']['worst'], d['config']['parity']['identical_path'])"
  done
done | tee gpurun_out/ab.log
cp /tmp/shipped.so cilqr_b200/lib/libcilqr_b200.so
