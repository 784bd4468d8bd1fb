#!/bin/bash
V=cilqr_b200/lib/variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 2 | tail -1
for v in g5 g6 g8; do timeout 300 python tools/occ_sweep.py --lib $V/libcilqr_b200_$v.so --horizon 100 --batch 65536 --pads 0 --reps 2 | tail -1; done
