#!/bin/bash
# Round 2, session D: multi-stage compaction relay sweep; DP planner ncu profile; new tests.
mkdir -p gpurun_out
for relay in "" "48,16,4" "64,24,6" "32,8,2" "48,12"; do
  for fl in 2 3; do
    echo "== relay '$relay' in-flight $fl"
    CILQR_B200_RELAY=$relay timeout 600 python bench.py --steps 6 --warmup 3 --in-flight $fl --no-cpu-baseline --no-corridor --no-dp --no-latency \
      > gpurun_out/r2d_bench.json 2>> gpurun_out/r2d_bench.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_bench.json"))
    print({k: round(d[k],1) for k in ("value","value_one_in_flight","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "launches", d["gpu_launches"])
except Exception as e:
    print("FAILED", e)
PY
  done
done 2>&1 | tee gpurun_out/r2d_sweep.log
tail -5 gpurun_out/r2d_bench.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_path_equals" 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dp_plan -c 1 -f -o gpurun_out/r2d_dp_prof \
    python tools/dp_bench.py --batch 2048 --base 1024 --reps 0 --cpu-sample 0 > gpurun_out/r2d_dp_ncu.log 2>&1; tail -3 gpurun_out/r2d_dp_ncu.log
ls -la gpurun_out | tail -5
