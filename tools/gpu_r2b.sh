#!/bin/bash
# Round 2, session B: strict-build bit parity, new parity tests, fp64 peak, updated bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_strict.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 | tee gpurun_out/r2b_pytest_strict.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dp.py -m gpu -q -s -k "gradient or 2048 or stage" 2>&1 | grep -v "^$" | tail -30 | tee gpurun_out/r2b_pytest_new.log
./tools/microbench/fp64_peak | tee gpurun_out/r2b_fp64_peak.json
python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json
