#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in 16 24 36; do CILQR_B200_CTX=$c timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 1; done
