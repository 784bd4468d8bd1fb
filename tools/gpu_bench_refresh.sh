#!/bin/bash
# the default bench line once more -> gpurun_out/r2zz_bench.json
mkdir -p gpurun_out
python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2zz_bench.json"))
print({k: round(d[k],1) for k in ("value","value_one_in_flight","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), d["e2e"]["steps"], "kernel_ms", round(d["roofline"]["kernel_ms"],2), "traffic", d["roofline"]["traffic"] is not None, "dp", round(d["dp_planner"]["traj_per_s"]))
print("parity", {k:v for k,v in d["config"]["parity"].items() if k in ("identical_path","within_1e-4")}, "latency", d["latency_b1"]["gpu_ms_p50"], d["clocks"])
PY
