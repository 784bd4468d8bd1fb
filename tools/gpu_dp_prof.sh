#!/bin/bash
# ncu --set full capture of one dp_plan_kernel launch (2048 scenes) -> gpurun_out/$1.ncu-rep
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dp_plan -s 1 -c 1 -f -o gpurun_out/$1 \
    python tools/dp_bench.py --batch 2048 --base 1024 --reps 1 --cpu-sample 0 > gpurun_out/$1.log 2>&1
tail -2 gpurun_out/$1.log; ls -la gpurun_out/$1.ncu-rep
