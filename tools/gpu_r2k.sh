#!/bin/bash
# Round 2, session K (2 GPUs): why is the NCCL all-gather at ~140 GB/s?  NCCL_DEBUG=INFO once, then channel-count knobs.
mkdir -p gpurun_out
run() {
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 \
     --no-corridor --no-dp --no-latency --no-cpu-baseline --no-e2e 2> gpurun_out/r2k_err_$1.log | grep '^{"metric' > gpurun_out/r2k_$1.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2k_$1.json"))
print("$1", round(d["value"]), round(d["value_one_in_flight"]), d["config"]["allgather"]["ms_per_step"], d["config"]["allgather"]["achieved_gbs_per_gpu"])
PY
}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH,TUNING run info
grep -E "NVLS|Channel|via P2P|via SHM|nChannels|Connected|Trees|Ring|transport|P2P" gpurun_out/r2k_err_info.log | head -40
NCCL_MIN_NCHANNELS=32 run minch32
NCCL_MIN_NCHANNELS=32 NCCL_P2P_NVL_CHUNKSIZE=1048576 NCCL_BUFFSIZE=16777216 run minch32_buf16m
nvidia-smi topo -m | head -8
