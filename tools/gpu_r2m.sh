#!/bin/bash
# Round 2, session M: artefacts of the final kernels -- ncu launch list of the bench command, ncu full capture of the solve
# kernel (DRAM traffic + FP64 operation counts -> profiles/traffic.json), horizon sweep (configs 1 and 4), memcheck.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-latency > gpurun_out/r2m_bench_ncu_launch.json 2> gpurun_out/r2m_err.log
timeout 900 ncu --set full --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum \
    --clock-control none --import-source on -k regex:cilqr_solve -c 1 -f -o gpurun_out/r2m_prof \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-corridor --no-dp --no-latency > gpurun_out/r2m_bench_ncu_full.json 2>> gpurun_out/r2m_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tracker_kernel -c 1 -f -o gpurun_out/r2m_tracker_prof \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-corridor --no-latency > /dev/null 2>> gpurun_out/r2m_err.log
timeout 600 python tools/horizon_sweep.py > gpurun_out/r2m_horizon_sweep.json 2>> gpurun_out/r2m_err.log; tail -c 900 gpurun_out/r2m_horizon_sweep.json
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import cilqr_b200 as cb
from cilqr_b200 import scenarios as sc
for N, B in ((20, 300), (100, 40)):
    batch = sc.generate(3, 0, B, N=N)
    s = cb.Solver(N_max=N, M_max=batch.M_max, S_max=batch.S, B_max=B)
    out = s.plan_batch(batch, trajectory=True, init_guess=True, hist_cap=4, result=True)
    U0 = out["init_controls"]
    o2 = s.plan_batch(batch, init_mode=1, init_controls=U0)
    o3 = s.plan_batch(batch, init_mode=2, init_states=out["init_states"], init_controls=U0)
    print(N, B, "converged", int((out["status"][:, 0] <= 2).sum()), "open-loop == default", bool(np.array_equal(o2["states"], out["states"])),
          "guess == default", bool(np.array_equal(o3["states"], out["states"])))
    s.close()
db = sc.generate_dp(5, 8, n_obs=6)
from oracle import dp_binding as dpo
s = cb.Solver()
dp = s.dp_plan_batch(db, dpo.build_barrier(db.ref))
good = dp["ok"].astype(bool) & np.isfinite(dp["trajectory"]).all(axis=(1, 2))
st4 = np.concatenate([db.start[good], np.full((good.sum(), 1), 10.0)], axis=1)
t = s.tracker_batch(st4, dp["trajectory"][good])
print("tracker ok", t["ok"].tolist())
s.close()
PY
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py 2>&1 | tail -12 | tee gpurun_out/r2m_sanitizer_memcheck.log
ls -la gpurun_out/r2m_*
