#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 300 python tools/corridor_bench.py --batch 16384 --reps 2 2>&1 | tail -2
timeout 300 python tools/corridor_bench.py --batch 65536 --reps 2 2>&1 | tail -2
