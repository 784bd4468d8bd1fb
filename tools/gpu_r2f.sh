#!/bin/bash
# Round 2, session F: nearest-segment certificate -- full GPU suite, bench, ncu launch list + full capture (traffic, FP64 ops).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "parity\]|strict\]|passed|failed|error" | tee gpurun_out/r2f_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print({k: round(d[k],1) for k in ("value","value_one_in_flight","ms_per_step")}, "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],2))
print("parity", {k:v for k,v in d["config"]["parity"].items() if k in ("identical_path","within_1e-4","worst")})
print("latency", d["latency_b1"])
print("dp", d["dp_planner"]["traj_per_s"], "corridor", d["corridor"]["traj_per_s"])
PY
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/r2f_prof \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-corridor --no-dp --no-latency > gpurun_out/r2f_bench_ncu_full.json 2>> gpurun_out/r2f_bench.err
ls -la gpurun_out/r2f_prof.ncu-rep
