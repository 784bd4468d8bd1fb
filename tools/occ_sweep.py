#!/usr/bin/env python
"""Development tool: throughput of the solve kernel vs resident warps per SM / library variant.

  python tools/occ_sweep.py --lib PATH --horizon 30 --batch 16384 --pads 0,8000,16000

Each (lib, pad) pair runs in its own process (the library is loaded once per process; the pad is
read from CILQR_B200_SMEM_PAD when the launch is planned).  Prints one line per configuration:
lib, N, warps/SM, smem/warp, ms per launch (best of 3, CUDA events inside the library), traj/s.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(a):
    import torch
    from cilqr_b200 import build as _build
    if a.lib:
        _build.LIB_PATH = os.path.abspath(a.lib)
        _build.is_stale = lambda: False
    import cilqr_b200
    from cilqr_b200 import scenarios
    N, B = a.horizon, a.batch
    batch = scenarios.generate(20260103, 0, B, N=N, n_obs=20, workers=16)
    dev = torch.device("cuda:0")
    tin = [torch.from_numpy(x).to(dev) for x in (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt,
                                                  batch.lane_left, batch.lane_right)]
    K = N + 1
    st = torch.empty((B, K, 6), dtype=torch.float64, device=dev)
    ct = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
    ss = torch.empty((B, 8), dtype=torch.float64, device=dev)
    solver = cilqr_b200.Solver(device=0, N_max=max(N, 100), M_max=batch.M_max, S_max=batch.S, B_max=B)
    ms = []
    for _ in range(a.reps + 1):
        solver.plan_batch_device(B, N, batch.M_max, batch.S, batch.S, *tin, st, ct, ss)
        solver.synchronize()
        ms.append(solver.last_kernel_ms())
    w, smem = solver.occupancy(N, batch.S, batch.S)
    try:
        print("stats", solver.debug_stats(), flush=True)
        hh = solver.completion_histogram()
        last = max(i for i, v in enumerate(hh) if v) + 1
        print("stats completions per 2 ms:", hh[:last], flush=True)
    except Exception as e:  # older variant libraries
        print("stats unavailable", e, flush=True)
    conv = int((ss[:, 0] <= 2).sum().item())
    best = min(ms[1:] or ms)
    print(json.dumps({"lib": os.path.basename(a.lib or "default"), "pad": int(os.environ.get("CILQR_B200_SMEM_PAD", "0")),
                      "N": N, "B": B, "warps_per_sm": w, "smem_per_warp": smem, "ms": round(best, 3),
                      "traj_per_s": round(conv / best * 1e3), "converged": conv,
                      "hash": int(ss[:, 7].sum().item()) % 1000003, "iters": float(ss[:, 1].mean().item())}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="")
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--batch", type=int, default=16384)
    ap.add_argument("--pads", default="0")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        child(a)
        return
    for pad in a.pads.split(","):
        env = dict(os.environ, CILQR_B200_SMEM_PAD=pad)
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--lib", a.lib, "--horizon", str(a.horizon),
               "--batch", str(a.batch), "--reps", str(a.reps)]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True)
        for l in r.stdout.splitlines():
            if l.startswith("stats"):
                print(l, flush=True)
        out = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(out[-1] if out else f"FAILED pad={pad}: {r.stderr[-400:]}", flush=True)


if __name__ == "__main__":
    main()
