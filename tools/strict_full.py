"""The strict build on ALL 65 536 scenarios of the bench workload against the oracle on the same portable libm."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cilqr_b200
from cilqr_b200 import scenarios
from oracle import binding_pm
binding_pm.build()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
batch = scenarios.generate(20260103, 0, B, N=100, workers=16)
s = cilqr_b200.Solver(device=0, N_max=100, M_max=batch.M_max, S_max=batch.S, B_max=B, variant="strict")
out = s.plan_batch(batch)
kms = s.last_kernel_ms()
s.close()
t = time.time()
Xo, Uo, So, _ = binding_pm.solve_batch(batch, nthreads=os.cpu_count() or 1)
cpu_s = time.time() - t
same = np.array([np.array_equal(out["states"][b], Xo[b], equal_nan=True) and np.array_equal(out["controls"][b], Uo[b], equal_nan=True)
                 and np.array_equal(out["status"][b], So[b], equal_nan=True) for b in range(B)])
print(f"[strict, full bench workload] B={B} N=100 seed 20260103: bit-identical scenarios {int(same.sum())}/{B} (states, controls and all "
      f"8 status words); strict kernel {kms:.0f} ms; oracle(pm libm) {cpu_s:.1f} s on {os.cpu_count()} host threads; exits "
      f"{np.bincount(So[:, 0].astype(int), minlength=5).tolist()}, iterations mean {So[:, 1].mean():.2f} max {int(So[:, 1].max())}")
sys.exit(0 if same.all() else 1)
