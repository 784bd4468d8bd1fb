#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_replay.py -m gpu -q -s -k "initial_guess or replay" 2>&1 | grep -E "init_mode|replay\]|passed|failed|rror|assert" | tee gpurun_out/r2h_pytest.log
