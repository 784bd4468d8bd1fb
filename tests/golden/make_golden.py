"""Generates tests/golden/cilqr_golden_v1.npz: a small frozen batch (inputs in the wire format of
include/cilqr_b200.h) together with the outputs of the CPU oracle (oracle/cilqr_oracle.c) on it.

The reference ships no fixtures and cannot be run here: these vectors pin the
oracle against regressions and give the GPU tests a committed target; they are NOT outputs of the
reference binary.  Re-run only when the oracle is deliberately changed:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cilqr_b200 import scenarios  # noqa: E402
from oracle import binding as oracle  # noqa: E402

if __name__ == "__main__":
    batch = scenarios.generate(424242, 0, 12, N=40, n_obs=11, M_max=12, S=24)
    X, U, S, conv = oracle.solve_batch(batch, nthreads=4)
    init = [oracle.solve(batch, b) for b in range(batch.B)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cilqr_golden_v1.npz")
    np.savez_compressed(path, N=batch.N, M_max=batch.M_max, S=batch.S, start=batch.start, coarse=batch.coarse,
                        corridor=batch.corridor, corridor_cnt=batch.corridor_cnt, lane_left=batch.lane_left,
                        lane_right=batch.lane_right, states=X, controls=U, status=S,
                        init_states=np.stack([r["init_states"] for r in init]),
                        cost_init=np.stack([r["cost_init"] for r in init]))
    print(path, os.path.getsize(path), "bytes; converged", conv, "/", batch.B, "iters", S[:, 1])
