#!/bin/bash
# compute-sanitizer memcheck of a small solve through the device and host paths (development aid)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import cilqr_b200 as cb
from cilqr_b200 import scenarios as sc
for N, B in ((20, 300), (100, 40)):
    batch = sc.generate(3, 0, B, N=N)
    s = cb.Solver(N_max=N, M_max=batch.M_max, S_max=batch.S, B_max=B)
    out = s.plan_batch(batch, trajectory=True, init_guess=True, hist_cap=4)
    print(N, B, "converged", int((out["status"][:, 0] <= 2).sum()), "stats", s.debug_stats())
    s.close()
PY
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py 2>&1 | tail -15 | tee gpurun_out/sanitizer_memcheck.log
