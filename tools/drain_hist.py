"""Completion histogram of one solve launch at the bench shape (2 ms buckets) + iteration-count distribution."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cilqr_b200
from cilqr_b200 import scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
batch = scenarios.generate(20260103, 0, B, N=100, workers=16)
dev = torch.device("cuda:0")
tin = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)]
X = torch.zeros(B, 101, 6, dtype=torch.float64, device=dev); U = torch.zeros(B, 100, 2, dtype=torch.float64, device=dev); S = torch.zeros(B, 8, dtype=torch.float64, device=dev)
s = cilqr_b200.Solver(device=0, N_max=100, M_max=batch.M_max, S_max=batch.S, B_max=B)
for _ in range(2):
    s.plan_batch_device(B, 100, batch.M_max, batch.S, batch.S, *tin, X, U, S)
    s.synchronize()
h = np.array(s.completion_histogram())
it = S[:, 1].cpu().numpy()
nz = np.nonzero(h)[0]
print(json.dumps({"kernel_ms": s.last_kernel_ms(), "hist_2ms": h[: nz[-1] + 1].tolist(),
                  "iters_quantiles": {q: float(np.quantile(it, q)) for q in (0.5, 0.9, 0.99, 0.999, 0.9999, 1.0)},
                  "iters_gt16": int((it > 16).sum()), "iters_gt32": int((it > 32).sum())}))
