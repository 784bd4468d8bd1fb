#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python tools/corridor_bench.py --batch 65536 --reps 2  2>&1 | tail -1 | tee gpurun_out/corridor_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corridor_build -c 1 -f -o gpurun_out/corr_prof \
    python tools/corridor_bench.py --batch 16384 --reps 0  > gpurun_out/corr_ncu.json 2> gpurun_out/corr_ncu.err; tail -2 gpurun_out/corr_ncu.err
