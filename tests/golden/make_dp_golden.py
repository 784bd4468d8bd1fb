"""Generates tests/golden/dp_golden_v1.npz: four DP-planner scenes (inputs in the layout of CilqrDpIn) with the
outputs of the CPU oracle (oracle/dp_oracle.c), which tests/test_dp_oracle_cross.py checks bit for bit against an
independent Python restatement.  The reference ships no fixtures for the planner (parity unpinned): these vectors
pin the oracle against regressions and give the GPU tests a committed target.
Re-run only deliberately:  python tests/golden/make_dp_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cilqr_b200 import scenarios  # noqa: E402
from oracle import dp_binding as dp  # noqa: E402

if __name__ == "__main__":
    db = scenarios.generate_dp(777, 4, n_obs=11)
    barrier = dp.build_barrier(db.ref)
    out = []
    for b in range(db.B):
        sc = dp.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                      db.dyn_poly[b], db.dyn_nv[b])
        out.append(dp.plan(sc, *db.start[b]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dp_golden_v1.npz")
    np.savez_compressed(path, ref=db.ref.astype(np.float64), barrier=barrier, start=db.start, static_poly=db.static_poly,
                        static_nv=db.static_nv, dyn_time=db.dyn_time, dyn_samples=db.dyn_samples, dyn_poly=db.dyn_poly,
                        dyn_nv=db.dyn_nv, ok=np.array([o[0] for o in out]), trajectory=np.stack([o[1] for o in out]),
                        cost=np.array([o[2] for o in out]), waypoints=np.stack([o[3] for o in out]))
    print(path, os.path.getsize(path), "bytes; ok", [o[0] for o in out], "cost", [round(o[2], 3) for o in out])
