#!/bin/bash
V=cilqr_b200/lib/variants
CILQR_B200_SMEM_PAD=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/v5_occ12 \
    python tools/occ_sweep.py --child --lib $V/libcilqr_b200_w16_b12.so --horizon 20 --batch 8192 --reps 0 2>&1 | tail -1
