#!/bin/bash
# Round 2, session ZZ (after the DP planner changes): final kernels -- full GPU suite, full bench line + reference arm, launch list, smoke().
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "parity\]|strict|grad exit|init_mode|replay\]|multi\]|tracker|dp parity|corridor|passed|failed|rror" | tee gpurun_out/r2zz_pytest_gpu.log
python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; tail -2 gpurun_out/r2zz_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2zz_bench_ref.json 2>> gpurun_out/r2zz_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2zz_bench.json"))
print({k: round(d[k],1) for k in ("value","value_one_in_flight","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],2), "launches", d["gpu_launches"])
print("parity", {k:v for k,v in d["config"]["parity"].items() if k in ("identical_path","within_1e-4","worst")}, d["config"]["parity"]["separation"]["strict_gpu_vs_oracle_pm_libm"]["bit_identical"])
print("latency", d["latency_b1"]["gpu_ms_p50"], d["latency_b1"]["gpu_ms_p99"], d["latency_b1"]["cpu_port_ms_p50"])
print("dp", d["dp_planner"]["traj_per_s"], "tracker", d["dp_planner"]["tracker"]["traj_per_s"], "corridor", d["corridor"]["traj_per_s"])
r=json.load(open("gpurun_out/r2zz_bench_ref.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2zz_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2zz_bench_ncu_launch.json 2>> gpurun_out/r2zz_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2zz_smoke.log
ls -la gpurun_out/r2zz_*
