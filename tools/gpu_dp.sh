#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12 | tee gpurun_out/pytest_gpu_dp.log
timeout 300 python tools/dp_bench.py --batch 8192 --base 512 --reps 1 2>&1 | tail -1 | tee gpurun_out/dp_bench.json
