#!/bin/bash
V=cilqr_b200/lib/variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 1
timeout 300 python tools/occ_sweep.py --lib $V/libcilqr_b200_c128.so --horizon 100 --batch 65536 --pads 0 --reps 1
CILQR_B200_CTX=96 timeout 300 python tools/occ_sweep.py --lib $V/libcilqr_b200_c128.so --horizon 100 --batch 65536 --pads 0 --reps 1
