#!/bin/bash
# One GPU session: GPU tests, bench line, ncu launch list, ncu full capture of the solve kernel at the bench config.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu_launch.json 2>> gpurun_out/bench.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_full.json 2>> gpurun_out/bench.err
ls -la gpurun_out
