#!/bin/bash
# Round 2, session E (2 GPUs): multi-GPU C ABI test, DP planner after the sample tables, full GPU suite, 2-GPU bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dp.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/r2e_pytest_multi_dp.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2e_pytest_gpu.log
timeout 300 python tools/dp_bench.py --batch 8192 --base 1024 --reps 2 --cpu-sample 0 | tee gpurun_out/r2e_dp_bench.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 \
   --no-corridor --no-dp --no-latency > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; tail -3 gpurun_out/r2e_bench_2gpu.err; cat gpurun_out/r2e_bench_2gpu.json
