#!/bin/bash
# A/B of library variants on the DP planner: every cilqr_b200/lib/variants/*.so is swapped in for libcilqr_b200.so in turn,
# tools/dp_bench.py times dp_plan_kernel (8192 scenes) and prints a checksum of the planned trajectories; then the DP parity
# suite (2048 scenes against the oracle) on the shipped build.  usage: bash tools/gpu_dp_ab.sh [rounds]
mkdir -p gpurun_out
cp cilqr_b200/lib/libcilqr_b200.so /tmp/shipped.so
for r in $(seq ${1:-2}); do
  for f in cilqr_b200/lib/variants/*.so; do
    v=$(basename $f .so)
    cp $f cilqr_b200/lib/libcilqr_b200.so
    python tools/dp_bench.py --batch 8192 --base 1024 --reps 2 --cpu-sample 2 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$v', round(d['traj_per_s']), [round(x,1) for x in d['ms']], d['planned_ok'], d.get('checksum'))"
  done
done | tee gpurun_out/dp_ab.log
cp /tmp/shipped.so cilqr_b200/lib/libcilqr_b200.so
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s 2>&1 | grep -E "dp parity|passed|failed" | tee -a gpurun_out/dp_ab.log
