#!/bin/bash
V=cilqr_b200/lib/variants
timeout 300 python tools/occ_sweep.py --lib $V/libcilqr_b200_c256.so --horizon 100 --batch 65536 --pads 0 --reps 1
timeout 300 python tools/occ_sweep.py --horizon 100 --batch 16384 --pads 0 --reps 1
timeout 300 python tools/occ_sweep.py --horizon 50 --batch 65536 --pads 0 --reps 1
timeout 300 python tools/occ_sweep.py --horizon 200 --batch 32768 --pads 0 --reps 1
