#!/bin/bash
# Round 2, session I (8 GPUs): BASELINE configs[3] as written -- 1 048 576 scenarios over 8 GPUs (131 072 per GPU), one all-gather.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 \
   --no-corridor --no-dp --no-latency > gpurun_out/r2i_bench_8gpu.json 2> gpurun_out/r2i_bench_8gpu.err; tail -5 gpurun_out/r2i_bench_8gpu.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2i_bench_8gpu.json"))
print({k: d[k] for k in ("value","value_one_in_flight","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
print(d["config"]["workload"], d["config"]["total_scenarios_per_step"], d["config"]["allgather"], "gen_s", d["config"]["scenario_gen_s"])
PY
