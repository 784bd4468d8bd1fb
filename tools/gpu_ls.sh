#!/bin/bash
for h in 16 10; do CILQR_B200_HOT=$h timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 1 | tail -3 | cut -c1-1200; done
