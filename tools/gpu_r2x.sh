#!/bin/bash
# A/B of the library variants, then the parity / strict / host-path suites on the shipped build
bash tools/gpu_ab.sh 2
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hostpath.py tests/test_gpu_strict.py -m gpu -x -q 2>&1 | tail -3
