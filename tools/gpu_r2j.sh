#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tracker.py -m gpu -q -s 2>&1 | grep -E "tracker|passed|failed|rror|assert" | tee gpurun_out/r2j_pytest_tracker.log
