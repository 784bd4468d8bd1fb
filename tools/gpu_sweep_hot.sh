#!/bin/bash
mkdir -p gpurun_out
for h in 6 10 16 24; do echo "hot_iter=$h"; CILQR_B200_HOT=$h timeout 300 python tools/occ_sweep.py --horizon 100 --batch 65536 --pads 0 --reps 2 | tail -1 | cut -c1-220; done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_corridor.py -m gpu -x -q -k "golden or edge or lane" 2>&1 | tail -6 | tee gpurun_out/sanitizer_corridor_memcheck.log
