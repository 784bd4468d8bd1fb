#!/bin/bash
# Round 2, session A: host-path fixes -- GPU tests, tests under CUDA_LAUNCH_BLOCKING=1, smoke() under ncu, short bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/r2a_pytest_gpu.log
CUDA_LAUNCH_BLOCKING=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_adapter.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2a_pytest_blocking.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2a_smoke_ncu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -3 gpurun_out/r2a_bench.err; cat gpurun_out/r2a_bench.json
