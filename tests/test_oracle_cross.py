"""The C oracle against the independent NumPy restatement (oracle/cilqr_numpy.py): two restatements
of the reference written in different styles must take the same decisions and agree to rounding.
This is the only pin available -- the reference has no tests, fixtures or golden vectors."""
import numpy as np

from cilqr_b200 import scenarios
from oracle import cilqr_numpy as cn


def test_stagewise_agreement(oracle):
    batch = scenarios.generate(11, 0, 2, N=20)
    for b in range(batch.B):
        c = oracle.Ctx(batch, b)
        s = cn.Solver(batch.start[b], batch.coarse[b], batch.corridor[b], batch.corridor_cnt[b],
                      batch.lane_left[b], batch.lane_right[b])
        X0, U0 = c.iqr()
        Xn, Un = s.iqr()
        np.testing.assert_allclose(X0, Xn, rtol=0, atol=1e-11)
        np.testing.assert_allclose(U0, Un, rtol=0, atol=1e-11)
        np.testing.assert_allclose(c.total_cost(X0, U0), s.total_cost(X0, U0), rtol=1e-12)
        lin = c.linearize(X0, U0)
        s.linearize(X0, U0)
        np.testing.assert_allclose(lin["A"], np.array(s.As), rtol=0, atol=1e-13)
        np.testing.assert_allclose(lin["B"], np.array(s.Bs), rtol=0, atol=1e-13)
        np.testing.assert_allclose(lin["Jx"], np.array(s.Jx), rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(lin["Hx"], np.array(s.Hx), rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(lin["Ju"], np.array(s.Ju), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(lin["Hu"], np.array(s.Hu), rtol=1e-12, atol=1e-12)
        for lam in (1.0, 1e-3):
            Ks, ks, dV = c.backward(lam)
            s.backward(lam)
            np.testing.assert_allclose(Ks, np.array(s.Ks), rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(ks, np.array(s.ks), rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(dV, s.dV, rtol=1e-8)
        Xf, Uf = c.forward(0.5012, X0, U0)
        Xg, Ug = s.forward(0.5012, X0, U0)
        np.testing.assert_allclose(Xf, Xg, rtol=0, atol=1e-10)
        np.testing.assert_allclose(Uf, Ug, rtol=0, atol=1e-10)
        c.close()


def test_full_solve_agreement(oracle):
    batch = scenarios.generate(7, 0, 3, N=30)
    for b in range(batch.B):
        r = cn.solve_scenario(batch, b)
        o = oracle.solve(batch, b, hist=True, trace=True)
        assert (r["status"], r["iters"]) == (o["status"], o["iters"])
        assert r["alphas"] == o["trace"][:, 1].astype(int).tolist()
        np.testing.assert_allclose(r["states"], o["states"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["controls"], o["controls"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["cost_hist"], o["cost_hist"], rtol=1e-8)
        np.testing.assert_allclose(r["init_states"], o["init_states"], rtol=0, atol=1e-11)


def test_backward_gains_solve_the_lq_subproblem(oracle):
    """Backward (ilqr_optimizer.cc:334-390) with lambda = 0 must return the minimiser of the
    quadratic model: checked against a dense KKT solve of the same LQ problem."""
    batch = scenarios.generate(5, 0, 1, N=8)
    c = oracle.Ctx(batch, 0)
    X0, U0 = c.iqr()
    lin = c.linearize(X0, U0)
    Ks, ks, _ = c.backward(0.0)
    N = batch.N
    # open-loop solution of min sum 1/2 dx'Hx dx + Jx'dx + 1/2 du'Hu du + Ju'du, dx+ = A dx + B du, dx0 = 0
    nz = N * 2
    G = np.zeros(((N + 1) * 6, nz))  # dx = G du
    for j in range(N):
        blk = lin["B"][j]
        for k in range(j + 1, N + 1):
            G[k * 6:(k + 1) * 6, j * 2:(j + 1) * 2] = blk
            if k < N:
                blk = lin["A"][k] @ blk
    Hx = np.zeros(((N + 1) * 6,) * 2)
    for k in range(N + 1):
        Hx[k * 6:(k + 1) * 6, k * 6:(k + 1) * 6] = lin["Hx"][k]
    Hu = np.zeros((nz, nz))
    for k in range(N):
        Hu[k * 2:(k + 1) * 2, k * 2:(k + 1) * 2] = lin["Hu"][k]
    H = G.T @ Hx @ G + Hu
    g = G.T @ lin["Jx"].reshape(-1) + lin["Ju"].reshape(-1)
    du = np.linalg.solve(H, -g)
    # closed-loop rollout of the gains on the linear model reproduces du
    dx = np.zeros(6)
    for k in range(N):
        duk = Ks[k] @ dx + ks[k]
        np.testing.assert_allclose(duk, du[k * 2:(k + 1) * 2], rtol=1e-6, atol=1e-8)
        dx = lin["A"][k] @ dx + lin["B"][k] @ duk
    c.close()
