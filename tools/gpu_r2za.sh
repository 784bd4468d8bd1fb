#!/bin/bash
# Round 2, session ZA (2 GPUs): full bench line at N=2 with the final kernel, and the multi-GPU C-ABI tests on two devices.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 4 --warmup 3 \
   --no-corridor --no-dp --no-latency 2> gpurun_out/r2za_bench_2gpu.err | grep '^{"metric' > gpurun_out/r2za_bench_2gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2za_bench_2gpu.json"))
print({k: d[k] for k in ("value","value_one_in_flight","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], d["config"]["allgather"])
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | grep -E "multi\]|passed|failed" | tee gpurun_out/r2za_pytest_multi.log
