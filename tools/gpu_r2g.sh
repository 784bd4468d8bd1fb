#!/bin/bash
# Round 2, session G: init modes (strict + production), replay tool, full GPU suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "init_mode|replay\]|strict, init|passed|failed|rror|assert" | tee gpurun_out/r2g_pytest_gpu.log
