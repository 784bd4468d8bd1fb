#!/bin/bash
mkdir -p gpurun_out
V=cilqr_b200/lib/variants
for cfg in "4 32000" "12 0"; do
  set -- $cfg
  CILQR_B200_SMEM_PAD=$2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/occ$1 \
    python tools/occ_sweep.py --child --lib $V/libcilqr_b200_w16_b12.so --horizon 20 --batch 8192 --reps 0 2>&1 | tail -2
done
# 1 and 2 warps per SM (timing only)
python tools/occ_sweep.py --lib $V/libcilqr_b200_w16_b12.so --horizon 20 --batch 8192 --pads 100000,60000,32000,0
ls -la gpurun_out
