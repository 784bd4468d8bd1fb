#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 2 | tail -2 | cut -c1-900
