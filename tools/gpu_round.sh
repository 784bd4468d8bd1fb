#!/bin/bash
# One GPU session: GPU tests, bench line (+ reference arm), ncu launch list, ncu full captures of the solve kernel
# and of the corridor build kernel at the bench shape.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_ncu_launch.json 2>> gpurun_out/bench.err
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-corridor --no-dp > gpurun_out/bench_ncu_full.json 2>> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corridor_build -c 1 -f -o gpurun_out/corr_prof \
    python tools/corridor_bench.py --batch 65536 --reps 0 > gpurun_out/corr_ncu.json 2>> gpurun_out/bench.err
timeout 300 python tools/horizon_sweep.py > gpurun_out/horizon_sweep.json 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/horizon_sweep.json
ls -la gpurun_out | tail -15
