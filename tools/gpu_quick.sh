#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-corridor > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; tail -5 gpurun_out/bench_pipe.err; cat gpurun_out/bench_pipe.json | cut -c1-1500
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-corridor --no-e2e --in-flight 3 2>&1 | tail -1 | cut -c1-400
