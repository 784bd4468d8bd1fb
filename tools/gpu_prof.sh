#!/bin/bash
# one ncu capture (--set full + FP64 op counters) of one solve launch at the bench config -> gpurun_out/$1.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/$1 \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-corridor --no-dp --no-latency > gpurun_out/$1_bench.json 2> gpurun_out/$1.err
ls -la gpurun_out/$1.ncu-rep
