#!/bin/bash
# Round 2, session C: drain relay -- correctness (full GPU suite incl. strict) and throughput with the relay on / off.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2c_pytest_gpu.log
for relay in "0,0" "16,2" "16,4" "32,2" "8,2"; do
  for fl in 2 3; do
    echo "== relay $relay in-flight $fl"
    CILQR_B200_RELAY=$relay timeout 600 python bench.py --steps 6 --warmup 3 --in-flight $fl --no-cpu-baseline --no-corridor --no-dp --no-latency \
      > gpurun_out/r2c_bench_${relay/,/_}_f$fl.json 2>> gpurun_out/r2c_bench.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_${relay/,/_}_f$fl.json"))
print({k: round(d[k],1) for k in ("value","value_one_in_flight","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "launches", d["gpu_launches"])
PY
  done
done 2>&1 | tee gpurun_out/r2c_sweep.log
tail -5 gpurun_out/r2c_bench.err
