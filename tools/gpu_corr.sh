#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python tools/corridor_bench.py --batch 65536 --reps 2  2>&1 | tail -1 | tee gpurun_out/corridor_bench.json
timeout 300 python tools/corridor_bench.py --batch 65536 --reps 2 --cap 88 2>&1 | tail -1
