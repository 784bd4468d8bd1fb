#!/bin/bash
CILQR_HOT_TIMING_DUMP=1 timeout 300 python tools/occ_sweep.py --child --lib cilqr_b200/lib/variants/libcilqr_b200_hott.so --horizon 100 --batch 65536 --reps 0 2>&1 | grep -E "hot phase|traj_per_s" | cut -c1-200
