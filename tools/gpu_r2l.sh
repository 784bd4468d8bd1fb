#!/bin/bash
# Round 2, session L: DP kernel after the code-size cut, full GPU suite, full bench line (+ reference arm), ncu launch list.
mkdir -p gpurun_out
timeout 300 python tools/dp_bench.py --batch 8192 --base 1024 --reps 2 --cpu-sample 0 | tee gpurun_out/r2l_dp_bench.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2l_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -3 gpurun_out/r2l_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2l_bench.json"))
print({k: round(d[k],1) for k in ("value","value_one_in_flight","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "launches", d["gpu_launches"])
print("dp", d["dp_planner"]["traj_per_s"], "tracker", d["dp_planner"]["tracker"], "corridor", d["corridor"]["traj_per_s"])
PY
