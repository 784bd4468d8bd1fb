#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for l in 16 24 32 48; do echo "hot_live=$l"; CILQR_B200_HOT_LIVE=$l timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 2 | tail -1 | cut -c1-200; done
