#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_strict.py tests/test_replay.py tests/test_gpu_multi.py tests/test_gpu_hostpath.py -m gpu -q -s 2>&1 | grep -E "init_mode|replay\]|strict, init|passed|failed|rror|assert" | tee gpurun_out/r2g_pytest_gpu2.log
