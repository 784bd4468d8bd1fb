"""Times the DP planner kernel: scenes generated for a base batch on the host and tiled on the device."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cilqr_b200  # noqa: E402
from cilqr_b200 import scenarios  # noqa: E402
from cilqr_b200.solver import dp_num_knots  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--base", type=int, default=1024)
    ap.add_argument("--obstacles", type=int, default=11)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=8)
    a = ap.parse_args()
    from oracle import dp_binding as dp
    dev = torch.device("cuda:0")
    db = scenarios.generate_dp(20260101, a.base, n_obs=a.obstacles)
    barrier = dp.build_barrier(db.ref)
    rep = max(1, a.batch // a.base)
    B, K = a.base * rep, dp_num_knots()
    one = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    tile = lambda x: one(x).repeat(rep, *([1] * (x.ndim - 1)))  # noqa: E731
    ref, bar = one(db.ref), one(barrier)
    tin = [tile(x) for x in (db.start, db.static_poly, db.static_nv, db.dyn_time, db.dyn_samples, db.dyn_poly, db.dyn_nv)]
    ok = torch.zeros(B, dtype=torch.int32, device=dev)
    coarse = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    s = cilqr_b200.Solver(device=0)
    ms = []
    for _ in range(a.reps + 1):
        s.dp_plan_batch_device(B, len(db.ref), len(barrier), 4, db.static_poly.shape[1], db.dyn_poly.shape[1],
                               db.dyn_poly.shape[2], ref, bar, tin[0], tin[1], tin[2], tin[3], tin[4], tin[5], tin[6], ok,
                               coarse=coarse)
        s.synchronize()
        ms.append(s.dp_last_kernel_ms())
    ms = ms[1:] or ms
    n = a.cpu_sample
    t0 = time.perf_counter()
    cfg = dp.default_config()
    okc = []
    for b in range(n):
        sc = dp.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                      db.dyn_poly[b], db.dyn_nv[b])
        okc.append(dp.plan(sc, *db.start[b], cfg)[0])
    cpu_s = (time.perf_counter() - t0) / max(n, 1)
    import hashlib
    checksum = hashlib.sha1(coarse.cpu().numpy().tobytes() + ok.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps({"B": B, "K": K, "obstacles": a.obstacles, "ms": ms, "traj_per_s": B / min(ms) * 1e3, "checksum": checksum,
                      "planned_ok": int(ok.sum().item()), "cpu_s_per_scene_1_thread": cpu_s,
                      "cpu_ok_equal": bool(np.array_equal(np.array(okc, bool), ok[:n].cpu().numpy().astype(bool)))}))
