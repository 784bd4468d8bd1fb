#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_zc.json 2> gpurun_out/bench_zc.err; tail -2 gpurun_out/bench_zc.err; python -c "
import json; b=json.load(open('gpurun_out/bench_zc.json')); print('value',b['value'],'e2e',b['e2e'])"
