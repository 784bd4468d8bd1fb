#!/bin/bash
# Round 2, session S: reciprocal micro-test; A/B of the working tree against libcilqr_b200_old.so (previous commit); GPU suites
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rcp_steps tools/microbench/rcp_steps.cu && /tmp/rcp_steps | tee gpurun_out/r2s_rcp_steps.json
bash tools/gpu_r2n.sh
