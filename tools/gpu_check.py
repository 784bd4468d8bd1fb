"""Quick GPU-vs-oracle diagnostic (development aid): stage dumps of the first iteration and full solves."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cilqr_b200 as cb
from cilqr_b200 import scenarios as sc
from oracle import binding as ob

np.set_printoptions(linewidth=200, precision=4)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
batch = sc.generate(7, 0, B, N=N)
K = N + 1
dev = torch.device("cuda:0")
solver = cb.Solver()
print("occupancy (warps/SM, smem/warp):", solver.occupancy(N, batch.S, batch.S))
tin = [torch.from_numpy(a).to(dev) for a in (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)]


def rel(a, b):
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0)))


# ---- stage dump on the first Bd scenarios
Bd = min(B, 16)
S2 = 2 * batch.S
z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
dbg = dict(corridor=z(Bd, K, batch.M_max, 3), lanes=z(Bd, S2, 3), X0=z(Bd, K, 6), U0=z(Bd, N, 2), cost0=z(Bd, 5),
           A11=z(Bd, N, 12), Jx=z(Bd, K, 6), Ju=z(Bd, N, 2), Hx=z(Bd, K, 9), Hu=z(Bd, N, 2), Kg=z(Bd, N, 12),
           kg=z(Bd, N, 2), dV=z(Bd, 2), Xn=z(Bd, K, 6), Un=z(Bd, N, 2), costn=z(Bd, 5),
           nearest=torch.zeros(Bd, K, 5, 2, dtype=torch.int32, device=dev))
tin_d = [t[:Bd].contiguous() for t in tin]
solver.debug_first_iteration(Bd, N, batch.M_max, batch.S, batch.S, *tin_d, dbg)
torch.cuda.synchronize()
g = {k: v.cpu().numpy() for k, v in dbg.items()}
worst = {}
for b in range(Bd):
    c = ob.Ctx(batch, b)
    cor, ll, lr = c.constraints()
    X0, U0 = c.iqr()
    c0 = c.total_cost(X0, U0)
    lin = c.linearize(X0, U0)
    Ks, ks, dV = c.backward(1.0)
    Xn, Un = c.forward(1.0, X0, U0)
    cn = c.total_cost(Xn, Un)
    A = lin["A"]; Bm = lin["B"]
    A11 = np.stack([A[:, 0, 2], A[:, 0, 3], A[:, 0, 4], A[:, 0, 5], A[:, 1, 2], A[:, 1, 3], A[:, 1, 4], A[:, 1, 5],
                    A[:, 2, 3], A[:, 2, 4], A[:, 2, 5], Bm[:, 2, 1]], axis=1)
    Hx = lin["Hx"]
    Hx9 = np.stack([Hx[:, 0, 0], Hx[:, 0, 1], Hx[:, 0, 2], Hx[:, 1, 1], Hx[:, 1, 2], Hx[:, 2, 2], Hx[:, 3, 3], Hx[:, 4, 4], Hx[:, 5, 5]], axis=1)
    Hu2 = np.stack([lin["Hu"][:, 0, 0], lin["Hu"][:, 1, 1]], axis=1)
    mask = (np.arange(batch.M_max)[None, :] < batch.corridor_cnt[b][:, None])
    errs = dict(corridor=rel(g["corridor"][b][mask], cor[mask]), lanes=rel(g["lanes"][b], np.concatenate([ll, lr])),
                X0=rel(g["X0"][b], X0), U0=rel(g["U0"][b], U0), cost0=rel(g["cost0"][b], c0), A11=rel(g["A11"][b], A11),
                Jx=rel(g["Jx"][b], lin["Jx"]), Ju=rel(g["Ju"][b], lin["Ju"]), Hx=rel(g["Hx"][b], Hx9), Hu=rel(g["Hu"][b], Hu2),
                Kg=rel(g["Kg"][b], Ks.reshape(N, 12)), kg=rel(g["kg"][b], ks), dV=rel(g["dV"][b], dV), Xn=rel(g["Xn"][b], Xn),
                Un=rel(g["Un"][b], Un), costn=rel(g["costn"][b], cn))
    for k_, v in errs.items():
        worst[k_] = max(worst.get(k_, 0.0), v)
print("stage max rel errors over", Bd, "scenarios:")
for k_, v in worst.items():
    print(f"  {k_:9s} {v:.3e}")

# ---- full solves
states = z(B, K, 6); controls = z(B, N, 2); status = z(B, 8)
for rep in range(2):
    torch.cuda.synchronize()
    t = time.time()
    solver.plan_batch_device(B, N, batch.M_max, batch.S, batch.S, *tin, states, controls, status)
    solver.synchronize()
    wall = time.time() - t
    print(f"device solve: B={B} wall {wall*1e3:.2f} ms kernel {solver.last_kernel_ms():.2f} ms -> {B/wall:.0f} traj/s")
t = time.time()
Xo, Uo, So, conv = ob.solve_batch(batch, nthreads=os.cpu_count())
print(f"oracle: {B/(time.time()-t):.1f} traj/s on {os.cpu_count()} threads, converged {conv}")
Xg, Ug, Sg = states.cpu().numpy(), controls.cpu().numpy(), status.cpu().numpy()
same_path = (Sg[:, 0] == So[:, 0]) & (Sg[:, 1] == So[:, 1]) & (Sg[:, 7] == So[:, 7])
ex = np.array([rel(Xg[b], Xo[b]) for b in range(B)])
eu = np.array([rel(Ug[b], Uo[b]) for b in range(B)])
print("decision path identical:", same_path.sum(), "/", B)
print("state err (same path): max %.3e  median %.3e" % (ex[same_path].max(), np.median(ex[same_path])))
print("control err (same path): max %.3e" % eu[same_path].max())
if (~same_path).any():
    idx = np.where(~same_path)[0][:10]
    print("different paths:", [(int(i), Sg[i, :2], So[i, :2], ex[i]) for i in idx])
print("status hist gpu", np.bincount(Sg[:, 0].astype(int), minlength=5), "iters mean", Sg[:, 1].mean())
# host path
out = solver.plan_batch(batch, trajectory=True, init_guess=True)
print("host path == device path:", np.array_equal(out["states"], Xg), np.array_equal(out["status"], Sg))
