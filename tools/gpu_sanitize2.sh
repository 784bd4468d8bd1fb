#!/bin/bash
# compute-sanitizer memcheck of the final kernels: a small solve (device + host path, result records, init modes), the DP planner,
# the tracker and the corridor builder (development aid; slow under the sanitizer, hence the small batches)
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import cilqr_b200 as cb
from cilqr_b200 import scenarios as sc
from cilqr_b200.solver import dp_num_knots
for N, B in ((20, 200), (100, 24)):
    batch = sc.generate(3, 0, B, N=N)
    s = cb.Solver(N_max=N, M_max=batch.M_max, S_max=batch.S, B_max=B)
    out = s.plan_batch(batch, trajectory=True, init_guess=True, hist_cap=4)
    print("solve", N, B, "converged", int((out["status"][:, 0] <= 2).sum()))
    s.close()
db = sc.generate_dp(11, 48, n_obs=11)
s = cb.Solver(device=0)
from oracle import dp_binding as dp   # only for the barrier builder of the test scenes
barrier = dp.build_barrier(db.ref)
o = s.dp_plan_batch(db, barrier) if hasattr(s, "dp_plan_batch") else None
print("dp", None if o is None else int(np.asarray(o["ok"]).sum()))
s.close()
PY
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san2.py 2>&1 | tail -15 | tee gpurun_out/r2_sanitizer_memcheck_final.log
