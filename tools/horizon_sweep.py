#!/usr/bin/env python
"""BASELINE.json configs[1] (1 024 scenarios, N=50) and configs[4] (horizon sweep N in {30,50,100,200} at batch
32 768): device-resident solve throughput and the HBM-roofline fraction on algorithmic bytes vs N.
usage: python tools/horizon_sweep.py > profiles/rNN_horizon_sweep.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cilqr_b200
from cilqr_b200 import scenarios

peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists("MEASURED_PEAKS.json") else 6650.0
dev = torch.device("cuda:0")
rows = []
for cfg, seed, B, N in (("configs[1]", 20260102, 1024, 50), ("configs[4]", 20260105, 32768, 30), ("configs[4]", 20260105, 32768, 50),
                        ("configs[4]", 20260105, 32768, 100), ("configs[4]", 20260105, 32768, 200)):
    batch = scenarios.generate(seed, 0, B, N=N, n_obs=20, workers=16)
    tin = [torch.from_numpy(x).to(dev) for x in (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt,
                                                  batch.lane_left, batch.lane_right)]
    K = N + 1
    st = torch.empty((B, K, 6), dtype=torch.float64, device=dev)
    ct = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
    ss = torch.empty((B, 8), dtype=torch.float64, device=dev)
    solver = cilqr_b200.Solver(device=0, N_max=N, M_max=batch.M_max, S_max=batch.S, B_max=B)
    ms = []
    for _ in range(4):
        solver.plan_batch_device(B, N, batch.M_max, batch.S, batch.S, *tin, st, ct, ss)
        solver.synchronize()
        ms.append(solver.last_kernel_ms())
    t = sorted(ms[1:])[1]
    conv = int((ss[:, 0] <= 2).sum().item())
    byt = scenarios.algorithmic_bytes(N, batch.M_max, batch.S)
    rows.append({"config": cfg, "batch": B, "N": N, "kernel_ms_median_of_3": round(t, 3), "converged": conv,
                 "traj_per_s": round(conv / t * 1e3), "mean_iterations": round(float(ss[:, 1].mean().item()), 3),
                 "algorithmic_bytes_per_traj": byt, "achieved_GBps": round(conv * byt / t / 1e6, 3),
                 "hbm_frac": round(conv * byt / t / 1e6 / peak, 6), "warps_per_sm": solver.occupancy(N, batch.S, batch.S)[0],
                 "smem_per_warp": solver.occupancy(N, batch.S, batch.S)[1]})
    solver.close()
    del tin, st, ct, ss
    torch.cuda.empty_cache()
print(json.dumps({"peak_hbm_GBps": peak, "rows": rows}, indent=1))
