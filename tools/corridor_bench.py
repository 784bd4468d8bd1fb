"""Times the corridor build kernel at the roofline-capture shape (B x K knots, n_obs obstacles):
scenarios are generated for a base batch on the host and tiled on the device."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cilqr_b200  # noqa: E402
from cilqr_b200 import scenarios  # noqa: E402
from cilqr_b200.solver import default_corridor_config  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--base", type=int, default=2048)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--obstacles", type=int, default=20)
    ap.add_argument("--mmax", type=int, default=20)
    ap.add_argument("--cap", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    _, ci = scenarios.generate_with_obstacles(20260103, 0, a.base, N=a.horizon, n_obs=a.obstacles)
    rep = a.batch // a.base
    B, K, P = a.base * rep, ci.K, ci.P_max
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev).repeat(rep, *([1] * (x.ndim - 1)))  # noqa: E731
    traj, pts, cnt = t(ci.traj), t(ci.obs_points), t(ci.obs_cnt)
    cor = torch.zeros(B, K, a.mmax, 3, dtype=torch.float64, device=dev)
    ccnt = torch.zeros(B, K, dtype=torch.int32, device=dev)
    code = torch.zeros(B, K, dtype=torch.int32, device=dev)
    s = cilqr_b200.Solver(device=0)
    cfg = default_corridor_config(point_cap=a.cap or 4 * a.obstacles + 16)
    ms = []
    for _ in range(a.reps + 1):
        s.corridor_batch_device(B, K, P, a.mmax, traj, pts, cnt, cor, ccnt, code, cfg=cfg)
        s.synchronize()
        ms.append(s.corridor_last_kernel_ms())
    ms = ms[1:] or ms
    codes = np.bincount(code.cpu().numpy().ravel(), minlength=6).tolist()
    in_bytes = traj.numel() * 8 + int(cnt.sum().item()) * 16 + cnt.numel() * 4
    out_bytes = int(ccnt.sum().item()) * 24 + ccnt.numel() * 8
    best = min(ms)
    print(json.dumps({"B": B, "K": K, "P_max": P, "point_cap": cfg.point_cap, "ms": ms, "knots_per_s": B * K / best * 1e3,
                      "traj_per_s": B / best * 1e3, "algorithmic_GBps": (in_bytes + out_bytes) / best / 1e6,
                      "planes_per_knot": float(ccnt.double().mean().item()), "codes": codes}))
