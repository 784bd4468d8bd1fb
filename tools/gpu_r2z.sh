#!/bin/bash
# Round 2, session Z (8 GPUs): BASELINE configs[3] with the final kernel -- 1 048 576 scenarios over 8 GPUs (131 072 per GPU), one all-gather.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 \
   --no-corridor --no-dp --no-latency > gpurun_out/r2z_bench_8gpu.out 2> gpurun_out/r2z_bench_8gpu.err; tail -3 gpurun_out/r2z_bench_8gpu.err
python - <<PY
import json
line=[l for l in open("gpurun_out/r2z_bench_8gpu.out") if l.startswith('{"metric')][-1]
open("gpurun_out/r2z_bench_8gpu.json","w").write(line)
d=json.loads(line)
print({k: d[k] for k in ("value","value_one_in_flight","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
print(d["config"]["workload"], d["config"]["total_scenarios_per_step"], d["config"]["allgather"], "gen_s", d["config"]["scenario_gen_s"])
PY
