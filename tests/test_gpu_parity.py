"""GPU parity: the PRODUCTION CUDA path (through the C ABI) against the CPU oracle on identical inputs.

What is compared with what.  The kernel's LOGIC is proven separately and exactly: the strict build of the same
kernel (reference-ordered arithmetic, portable libm) reproduces the oracle bit for bit on every scenario
(tests/test_gpu_strict.py).  The production build evaluates the same expressions with re-associated sums (warp
reductions over knots, log of a product instead of a sum of logs), fused multiply-adds and CUDA's libm, i.e. it
differs from the reference by ~1e-16 relative per operation.  A CILQR solve is a chain of accept / reject decisions
with a stiff barrier; on a few ill-conditioned scenarios it amplifies such rounding by many orders of magnitude.
That tail is a property of the reference algorithm, not of this kernel: the oracle shows the same tail against
ITSELF when only its libm is exchanged (glibc -> pm_math.h: 2043/2048 identical paths, 8 of those beyond 1e-4) or only
FMA contraction is enabled (2044/2048, 13 beyond 1e-4) -- tests/test_oracle_golden.py.  Hence, for the production build:
  * stage level (one linearisation / backward / forward / cost on the same iterate): 1e-9 relative;
  * full solves: identical (status, iteration count, line-search sequence) and, on those, states/controls within
    1e-4 relative (the north-star tolerance) on all but a measured tail.  Measured on B200 (profiles/r02_*):
    N=30 256/256 and 1.0000; N=50 1023/1024 and 0.9990; N=100 510/512 and 0.9980; N=200 64/64 and 1.0000;
    shipped road N=80 127/128 and 0.9921; the 4 096-scenario sample of the bench workload 4085/4096 and 0.9973.
    The thresholds below are those figures minus a small margin; every mismatch is counted and printed, never hidden.
"""
import ctypes as C

import numpy as np
import pytest

from cilqr_b200 import scenarios

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.abs(a - b) / (np.abs(b) + 1.0)


def _dev(batch, torch, dev):
    return [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
            (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)]


def _solve_device(solver, batch):
    import torch
    dev = torch.device("cuda:0")
    B, N, K = batch.B, batch.N, batch.N + 1
    tin = _dev(batch, torch, dev)
    X = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    U = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    S = torch.zeros(B, 8, dtype=torch.float64, device=dev)
    solver.plan_batch_device(B, N, batch.M_max, batch.lane_left.shape[1], batch.lane_right.shape[1], *tin, X, U, S)
    solver.synchronize()
    return X.cpu().numpy(), U.cpu().numpy(), S.cpu().numpy()


def _compare(oracle, batch, Xg, Ug, Sg, min_same=0.985, min_within=0.985):
    import os
    Xo, Uo, So, _ = oracle.solve_batch(batch, nthreads=os.cpu_count() or 1)
    same = (Sg[:, 0] == So[:, 0]) & (Sg[:, 1] == So[:, 1]) & (Sg[:, 7] == So[:, 7])
    ex = rel(Xg, Xo).reshape(batch.B, -1).max(axis=1)
    eu = rel(Ug, Uo).reshape(batch.B, -1).max(axis=1)
    e = np.maximum(ex, eu)
    print(f"\n[parity] B={batch.B} N={batch.N}: identical decision path {same.sum()}/{batch.B}; on those: "
          f"median {np.median(e[same]):.2e} p99 {np.quantile(e[same], 0.99):.2e} max {e[same].max():.2e}; "
          f"within 1e-4: {(e[same] < 1e-4).mean():.4f}; different-path scenarios: {np.where(~same)[0][:16].tolist()}")
    assert same.mean() >= min_same
    assert np.median(e[same]) < 1e-10
    assert (e[same] < 1e-4).mean() >= min_within
    # cost of the returned trajectory (status record) against the oracle's for identical paths
    assert np.quantile(rel(Sg[same, 2:7], So[same, 2:7]).max(axis=1), 0.99) < 1e-6
    return same, e


def test_stage_level_parity(solver, oracle):
    """Every stage of the first iteration on the same iterate (cilqr_debug_first_iteration):
    constraints (a4), iqr (a5), TotalCost (a15), A/B + cost derivatives (a7,a9,a10), Backward (a12)."""
    import torch
    dev = torch.device("cuda:0")
    batch = scenarios.generate(7, 0, 24, N=100)
    B, N, K, S2 = batch.B, batch.N, batch.N + 1, 2 * batch.S
    z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)  # noqa: E731
    dbg = dict(corridor=z(B, K, batch.M_max, 3), lanes=z(B, S2, 3), X0=z(B, K, 6), U0=z(B, N, 2), cost0=z(B, 5),
               A11=z(B, N, 12), Jx=z(B, K, 6), Ju=z(B, N, 2), Hx=z(B, K, 9), Hu=z(B, N, 2), Kg=z(B, N, 12),
               kg=z(B, N, 2), dV=z(B, 2), Xn=z(B, K, 6), Un=z(B, N, 2), costn=z(B, 5),
               nearest=torch.zeros(B, K, 5, 2, dtype=torch.int32, device=dev), gnorm=z(B))
    solver.debug_first_iteration(B, N, batch.M_max, batch.S, batch.S, *_dev(batch, torch, dev), dbg)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in dbg.items()}
    worst = {}
    for b in range(B):
        c = oracle.Ctx(batch, b)
        cor, ll, lr = c.constraints()
        X0, U0 = c.iqr()
        lin = c.linearize(X0, U0)
        Ks, ks, dV = c.backward(1.0)
        A, Bm, Hx = lin["A"], lin["B"], lin["Hx"]
        A11 = np.stack([A[:, 0, 2], A[:, 0, 3], A[:, 0, 4], A[:, 0, 5], A[:, 1, 2], A[:, 1, 3], A[:, 1, 4], A[:, 1, 5],
                        A[:, 2, 3], A[:, 2, 4], A[:, 2, 5], Bm[:, 2, 1]], axis=1)
        Hx9 = np.stack([Hx[:, 0, 0], Hx[:, 0, 1], Hx[:, 0, 2], Hx[:, 1, 1], Hx[:, 1, 2], Hx[:, 2, 2], Hx[:, 3, 3],
                        Hx[:, 4, 4], Hx[:, 5, 5]], axis=1)
        Hu2 = np.stack([lin["Hu"][:, 0, 0], lin["Hu"][:, 1, 1]], axis=1)
        # structure the kernel relies on: everything outside the stored entries is exactly zero / identity
        assert np.all(lin["Hu"][:, 0, 1] == 0) and np.all(Hx[:, 3:, :3] == 0) and np.all(Hx[:, 3, 4:] == 0)
        mask = np.arange(batch.M_max)[None, :] < batch.corridor_cnt[b][:, None]
        # the gains of a stiff problem are compared relative to the row scale
        errs = dict(corridor=rel(g["corridor"][b][mask], cor[mask]).max(),
                    lanes=rel(g["lanes"][b], np.concatenate([ll, lr])).max(),
                    X0=rel(g["X0"][b], X0).max(), U0=rel(g["U0"][b], U0).max(),
                    cost0=rel(g["cost0"][b], c.total_cost(X0, U0)).max(), A11=rel(g["A11"][b], A11).max(),
                    Jx=(np.abs(g["Jx"][b] - lin["Jx"]) / (np.abs(lin["Jx"]).max() + 1)).max(),
                    Ju=rel(g["Ju"][b], lin["Ju"]).max(),
                    Hx=(np.abs(g["Hx"][b] - Hx9) / (np.abs(Hx9).max() + 1)).max(), Hu=rel(g["Hu"][b], Hu2).max(),
                    Kg=(np.abs(g["Kg"][b] - Ks.reshape(N, 12)) / (np.abs(Ks).max() + 1)).max(),
                    kg=(np.abs(g["kg"][b] - ks) / (np.abs(ks).max() + 1)).max(), dV=rel(g["dV"][b], dV).max(),
                    # CalGradientNorm (ilqr_optimizer.cc:322-332) of those gains
                    gnorm=rel(g["gnorm"][b], np.mean(np.max(np.abs(ks) / (np.abs(U0) + 1.0), axis=1))))
        # nearest lane segment indices are integers: exact
        near = np.array([[[c.nearest(side, X0[k, 0] + off * np.cos(X0[k, 2]), X0[k, 1] + off * np.sin(X0[k, 2]))
                           for side in (0, 1)] for off in _disc_offsets()] for k in range(K)])
        assert np.array_equal(g["nearest"][b], near)
        for k_, v in errs.items():
            worst[k_] = max(worst.get(k_, 0.0), float(v))
        c.close()
    print("\n[stage] max relative errors:", {k: f"{v:.1e}" for k, v in worst.items()})
    for k_, v in worst.items():
        assert v < 1e-9, (k_, v)


def _disc_offsets():
    Ld = (0.929 + 1.0 + 0.96) / 5
    return [Ld * (j - 0.5) - 0.929 for j in range(5)]


@pytest.mark.parametrize("N,B,seed", [(30, 256, 11), (50, 1024, 20260102), (100, 512, 7), (200, 64, 13)])
def test_full_solve_parity(solver, oracle, N, B, seed):
    """configs[1] (1 024 x N=50) and slices of configs[2] / the horizon sweep {30, 50, 100, 200}."""
    batch = scenarios.generate(seed, 0, B, N=N)
    Xg, Ug, Sg = _solve_device(solver, batch)
    _compare(oracle, batch, Xg, Ug, Sg)


def test_gradient_norm_exit(oracle):
    """The gradient-norm exit (ilqr_optimizer.cc:235-241): scenarios that leave through it when both cost
    tolerances are 0 -- the same ones tests/test_reference_pins.py pins reference <-> oracle bit for bit.  The GPU
    must take that exit too, after the same accepts, with the trajectory of the last accept."""
    import cilqr_b200
    from picks import GRAD_EXIT_PICKS
    p = cilqr_b200.solver.default_params()
    p.abs_cost_tol = p.rel_cost_tol = 0.0
    po = oracle.default_params()
    po.abs_cost_tol = po.rel_cost_tol = 0.0
    s = cilqr_b200.Solver(params=p, device=0)
    n_grad = n_same = 0
    for (seed, N), ids in GRAD_EXIT_PICKS.items():
        full = scenarios.generate(seed, 0, max(ids) + 1, N=N)
        batch = scenarios.ScenarioBatch(full.N, full.M_max, full.S, *[np.ascontiguousarray(a[ids]) for a in
                                        (full.start, full.coarse, full.corridor, full.corridor_cnt, full.lane_left,
                                         full.lane_right)])
        out = s.plan_batch(batch)
        Xo, Uo, So, _ = oracle.solve_batch(batch, params=po, nthreads=4)
        assert (So[:, 0] == 2).all()
        same = (out["status"][:, 0] == So[:, 0]) & (out["status"][:, 1] == So[:, 1]) & (out["status"][:, 7] == So[:, 7])
        e = np.maximum(rel(out["states"], Xo).reshape(len(ids), -1).max(axis=1),
                       rel(out["controls"], Uo).reshape(len(ids), -1).max(axis=1))
        print(f"\n[grad exit] seed {seed} N={N}: GPU exits {out['status'][:, 0].astype(int).tolist()} iterations "
              f"{out['status'][:, 1].astype(int).tolist()} (oracle {So[:, 1].astype(int).tolist()}); identical path "
              f"{int(same.sum())}/{len(ids)}; max rel err on those {e[same].max() if same.any() else float('nan'):.2e}")
        n_grad += int((out["status"][:, 0] == 2).sum())
        n_same += int(same.sum())
        assert e[same].max() < 1e-6 if same.any() else True
    s.close()
    # 10-49 iterations at lambda -> 1e-8 amplify rounding differences; the exit itself must be reached on most
    assert n_grad >= 6 and n_same >= 5, (n_grad, n_same)


def test_initial_guess_modes(solver, oracle):
    """init_mode of CilqrBatchIn on the production build (ilqr_optimizer.cc:168-169).  (a) iqr's own result handed back
    as the caller's guess (mode 2) or as controls to roll out open loop (mode 1) is the default solve again: same
    parity as the default path.  (b) a rough guess (noise controls: ~14 iterations from a cost of 1e5-1e7) is reported,
    not thresholded tightly: the longer and stiffer the solve, the more rounding it amplifies -- the strict build
    reproduces the oracle bit for bit on exactly these inputs (tests/test_gpu_strict.py)."""
    import os
    batch = scenarios.generate(23, 0, 256, N=50)
    nt = os.cpu_count() or 1
    ref = solver.plan_batch(batch, init_guess=True)
    X0, U0 = ref["init_states"], ref["init_controls"]
    for mode, kw in ((1, dict(init_controls=U0)), (2, dict(init_states=X0, init_controls=U0))):
        out = solver.plan_batch(batch, init_mode=mode, **kw)
        same_as_default = np.array([np.array_equal(out["states"][b], ref["states"][b]) and np.array_equal(out["status"][b], ref["status"][b])
                                    for b in range(batch.B)])
        Xo, Uo, So, _ = oracle.solve_batch(batch, nthreads=nt, init_mode=mode, **kw)
        same = (out["status"][:, 0] == So[:, 0]) & (out["status"][:, 1] == So[:, 1]) & (out["status"][:, 7] == So[:, 7])
        e = np.maximum(rel(out["states"], Xo).reshape(batch.B, -1).max(axis=1), rel(out["controls"], Uo).reshape(batch.B, -1).max(axis=1))
        print(f"\n[init_mode {mode}, iqr's own guess] bit-identical to the default solve {same_as_default.sum()}/{batch.B}; vs oracle: "
              f"identical path {same.sum()}/{batch.B}, within 1e-4 on those {(e[same] < 1e-4).mean():.4f}")
        assert same_as_default.mean() >= 0.98  # (mode 1 re-rolls the guess: -0.0 controls may become +0.0)
        assert same.mean() >= 0.98 and (e[same] < 1e-4).mean() >= 0.98
    rng = np.random.default_rng(0)
    Xg = batch.coarse.copy()
    Xg[:, 0, :4] = batch.start
    Xg[:, 0, 4:] = 0.0
    Ug = rng.normal(0.0, 0.05, size=(batch.B, batch.N, 2))
    for mode, kw in ((1, dict(init_controls=Ug)), (2, dict(init_states=Xg, init_controls=Ug))):
        out = solver.plan_batch(batch, init_mode=mode, **kw)
        Xo, Uo, So, _ = oracle.solve_batch(batch, nthreads=nt, init_mode=mode, **kw)
        same = (out["status"][:, 0] == So[:, 0]) & (out["status"][:, 1] == So[:, 1]) & (out["status"][:, 7] == So[:, 7])
        e = np.maximum(rel(out["states"], Xo).reshape(batch.B, -1).max(axis=1), rel(out["controls"], Uo).reshape(batch.B, -1).max(axis=1))
        print(f"\n[init_mode {mode}, rough guess] identical decision path {same.sum()}/{batch.B}; within 1e-4 on those "
              f"{(e[same] < 1e-4).mean():.4f}; median {np.median(e[same]):.2e}; mean iterations {So[:, 1].mean():.2f}; exits "
              f"{np.bincount(So[:, 0].astype(int), minlength=5).tolist()}")
        assert same.mean() >= 0.6 and np.median(e[same]) < 1e-6  # measured: 212/256 (mode 1), see the docstring
        assert (out["status"][:, 0] == So[:, 0]).mean() >= 0.9   # same exit flag


def test_shipped_road_and_horizon(solver, oracle):
    """configs[0] emulated: N = 80, 11 obstacles, the shipped (tight) road radii; many of these end on
    the relative-cost test after 0-2 iterations with a huge cost -- same exits as the oracle."""
    batch = scenarios.generate(20260101, 0, 128, N=80, n_obs=11, road_name="shipped")
    Xg, Ug, Sg = _solve_device(solver, batch)
    _compare(oracle, batch, Xg, Ug, Sg, min_same=0.97, min_within=0.97)


def test_golden_fixture(solver):
    """Committed fixture (tests/golden/make_golden.py): inputs + oracle outputs."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cilqr_golden_v1.npz"))
    batch = scenarios.ScenarioBatch(int(z["N"]), int(z["M_max"]), int(z["S"]), z["start"], z["coarse"], z["corridor"],
                                    z["corridor_cnt"], z["lane_left"], z["lane_right"])
    Xg, Ug, Sg = _solve_device(solver, batch)
    same = (Sg[:, 0] == z["status"][:, 0]) & (Sg[:, 1] == z["status"][:, 1]) & (Sg[:, 7] == z["status"][:, 7])
    assert same.sum() >= batch.B - 1
    assert rel(Xg[same], z["states"][same]).max() < 1e-6
    assert rel(Ug[same], z["controls"][same]).max() < 1e-6


def test_host_path_equals_device_path_and_is_deterministic(solver):
    batch = scenarios.generate(3, 0, 300, N=50)
    X1, U1, S1 = _solve_device(solver, batch)
    X2, U2, S2 = _solve_device(solver, batch)
    assert np.array_equal(X1, X2) and np.array_equal(U1, U2) and np.array_equal(S1, S2)  # bitwise
    out = solver.plan_batch(batch, trajectory=True, init_guess=True, result=True)
    assert np.array_equal(out["states"], X1) and np.array_equal(out["controls"], U1) and np.array_equal(out["status"], S1)
    # TrajectoryPlanner::Plan's post-processing (trajectory_planner.cpp:103-125): s = running sum of hypot, the
    # rest as TransformToTrajectory
    res, trj = out["result"], out["trajectory"]
    seg = np.hypot(np.diff(X1[:, :, 0], axis=1), np.diff(X1[:, :, 1], axis=1))
    s_ref = np.concatenate([np.zeros((batch.B, 1)), np.cumsum(seg, axis=1)], axis=1)  # sequential, like :110-112
    np.testing.assert_allclose(res[:, :, 1], s_ref, rtol=1e-14, atol=0)
    assert np.array_equal(np.delete(res, 1, axis=2), np.delete(trj, 1, axis=2)) and np.all(trj[:, :, 1] == 0)
    # order independence: scenario b's result does not depend on its position in the batch
    perm = np.random.default_rng(0).permutation(batch.B)
    pb = scenarios.ScenarioBatch(batch.N, batch.M_max, batch.S, *[np.ascontiguousarray(a[perm]) for a in
                                 (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)])
    Xp, Up, Sp = _solve_device(solver, pb)
    assert np.array_equal(Xp, X1[perm]) and np.array_equal(Sp, S1[perm])
    # TransformToTrajectory record (ilqr_optimizer.cc:771-791)
    tr = out["trajectory"]
    assert np.array_equal(tr[:, :, [2, 3, 4, 6, 7, 9]], X1)
    np.testing.assert_allclose(tr[:, :, 5], np.tan(X1[:, :, 5]), rtol=1e-14)
    assert np.array_equal(tr[:, :-1, [8, 10]], U1) and np.all(tr[:, -1, [8, 10]] == 0)
    np.testing.assert_allclose(tr[:, :, 0], np.broadcast_to(np.arange(batch.N + 1) * 0.1, tr[:, :, 0].shape), atol=1e-12)


def test_host_path_pinned_outputs_are_written_in_place(solver):
    """Outputs in pinned (device-mapped) host memory are written by the kernel directly (no D2H pass);
    pageable outputs go through device staging.  Both must be bit-identical, with pinned inputs too, and
    with a batch that spans several H2D chunks of the streaming host path."""
    import torch
    batch = scenarios.generate(9, 0, 700, N=30)
    ref = solver.plan_batch(batch, trajectory=True, init_guess=True, hist_cap=3)
    K, N, B = batch.N + 1, batch.N, batch.B
    pin = lambda *shape, dt=torch.float64: torch.zeros(shape, dtype=dt).pin_memory().numpy()  # noqa: E731
    out = dict(states=pin(B, K, 6), controls=pin(B, N, 2), status=pin(B, 8), trajectory=pin(B, K, 13),
               init_states=pin(B, K, 6), init_controls=pin(B, N, 2), cost_hist=pin(B, 3, 5),
               iter_states=pin(B, 3, K, 6), iter_controls=pin(B, 3, N, 2), hist_len=pin(B, 2, dt=torch.int32))
    pb = scenarios.ScenarioBatch(batch.N, batch.M_max, batch.S, *[torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
                                 for a in (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt,
                                           batch.lane_left, batch.lane_right)])
    import os
    os.environ["CILQR_CHUNK"] = "128"  # read at handle creation: several chunks behind the watermark
    try:
        import cilqr_b200
        s2 = cilqr_b200.Solver(device=0, N_max=N, M_max=batch.M_max, S_max=batch.S, B_max=B)
    finally:
        del os.environ["CILQR_CHUNK"]
    got = s2.plan_batch(pb, trajectory=True, init_guess=True, hist_cap=3, out=out)
    s2.close()
    for k, v in ref.items():
        assert np.array_equal(got[k], v), k


def test_history_outputs(solver, oracle):
    """cost_ (ilqr_optimizer.cc:173,283,296) and iter_trajs (:170,294) histories, initial guess."""
    batch = scenarios.generate(9, 0, 16, N=40)
    out = solver.plan_batch(batch, init_guess=True, hist_cap=64)
    for b in range(batch.B):
        o = oracle.solve(batch, b, hist=True)
        if (out["status"][b, 0], out["status"][b, 1]) != (o["status"], o["iters"]):
            continue
        n_cost, n_it = out["hist_len"][b]
        assert n_cost == len(o["cost_hist"])
        np.testing.assert_allclose(out["cost_hist"][b, :n_cost], o["cost_hist"], rtol=1e-7)
        assert n_it == n_cost - (1 if o["status"] <= 1 else 0)
        np.testing.assert_allclose(out["init_states"][b], o["init_states"], rtol=0, atol=1e-9)
        np.testing.assert_array_equal(out["iter_states"][b, 0], out["init_states"][b])
        np.testing.assert_array_equal(out["iter_controls"][b, 0], out["init_controls"][b])


def test_size_independent_properties_at_full_size(solver):
    """configs[2] shape (N = 100, M_max = 20) at a batch far beyond what the oracle checks in seconds:
    properties every output of the reference algorithm has, verified on all scenarios.
      * x_{k+1} = Dynamics(x_k, u_k) (midpoint RK2, vehicle_model.cc:88-121) from x_0 = (start, 0, 0);
      * theta, delta wrapped into [-pi, pi); delta_rate wrapped (ilqr_optimizer.cc:408);
      * status in {0..4}, iteration count <= max_iter_num, total cost = sum of its four parts;
      * the returned cost never exceeds the cost of the initial guess (only improving steps are accepted)."""
    import torch
    batch = scenarios.generate(20260103, 0, 16384, N=100)
    out = solver.plan_batch(batch, init_guess=True)
    X, U, S = (torch.from_numpy(out[k]) for k in ("states", "controls", "status"))
    assert torch.isfinite(X).all() and torch.isfinite(U).all() and torch.isfinite(S).all()
    x = X[:, 0].clone()
    st = torch.from_numpy(batch.start)
    assert torch.equal(x[:, :4], st) and torch.all(x[:, 4:] == 0)
    dt, L = 0.1, 1.0

    def wrap(a):
        r = torch.fmod(a + np.pi, 2 * np.pi)
        r = torch.where(r < 0, r + 2 * np.pi, r)
        return r - np.pi

    def f(s, u):
        th, v, a, de = wrap(s[:, 2]), s[:, 3], s[:, 4], wrap(s[:, 5])
        return torch.stack([v * torch.cos(th), v * torch.sin(th), v * torch.tan(de) / L, a, u[:, 0], u[:, 1]], dim=1)

    worst = 0.0
    for k in range(batch.N):
        k1 = f(X[:, k], U[:, k])
        k2 = f(X[:, k] + 0.5 * dt * k1, U[:, k])
        nx = X[:, k] + dt * k2
        nx[:, 2], nx[:, 5] = wrap(nx[:, 2]), wrap(nx[:, 5])
        worst = max(worst, float(((nx - X[:, k + 1]).abs() / (X[:, k + 1].abs() + 1)).max()))
    # (the production build contracts x + dt * k into one fused multiply-add; a diverging trial step of a scenario that ends
    # in the lambda-overflow exit has |dt * k| ~ 1e4 in the heading, where half an ulp is 1e-12 -- and the wrap into
    # [-pi, pi) turns that into an absolute error of the O(1) result.  The CPU restatement, unfused, checks to 1.5e-16.)
    print(f"\n[properties] B={batch.B} N={batch.N}: worst one-step dynamics residual {worst:.2e}")
    assert worst < 1e-10, worst
    assert (X[:, :, 2] >= -np.pi).all() and (X[:, :, 2] < np.pi).all()
    assert (U[:, :, 1] >= -np.pi).all() and (U[:, :, 1] < np.pi).all()
    assert set(S[:, 0].long().unique().tolist()) <= {0, 1, 2, 3, 4}
    assert (S[:, 1] <= 200).all() and (S[:, 1] >= 0).all()
    assert torch.allclose(S[:, 2], S[:, 3:7].sum(dim=1), rtol=1e-13, atol=0)
    # checksum of checksums: re-solving the two halves separately gives the same per-scenario status records
    half = batch.B // 2
    o1 = solver.plan_batch(batch.slice(0, half))
    o2 = solver.plan_batch(batch.slice(half, batch.B))
    assert np.array_equal(np.concatenate([o1["status"], o2["status"]]), out["status"])


def test_edge_cases(solver, oracle):
    import cilqr_b200
    # ragged constraint sets: zero planes at some knots, a single plane at others, one lane segment on one side
    batch = scenarios.generate(21, 0, 8, N=20)
    batch.corridor_cnt[:, ::3] = 0
    batch.corridor_cnt[:, 1::3] = 1
    rb = scenarios.ScenarioBatch(batch.N, batch.M_max, batch.S, batch.start, batch.coarse, batch.corridor,
                                 batch.corridor_cnt, np.ascontiguousarray(batch.lane_left[:, 3:4]), batch.lane_right)
    Xg, Ug, Sg = _solve_device(solver, rb)
    _compare(oracle, rb, Xg, Ug, Sg, min_same=0.875)  # measured 8/8
    # shortest horizon
    b1 = scenarios.generate(22, 0, 4, N=1)
    Xg, Ug, Sg = _solve_device(solver, b1)
    _compare(oracle, b1, Xg, Ug, Sg, min_same=0.75)  # measured 4/4
    # empty batch is a no-op
    e = batch.slice(0, 0)
    out = solver.plan_batch(e)
    assert out["states"].shape == (0, 21, 6)
    # guards of IlqrOptimizer::Plan (ilqr_optimizer.cc:64-78) -> CILQR_E_INVALID; capacity -> CILQR_E_CAPACITY
    L = cilqr_b200.load_library()
    bi = cilqr_b200.solver.BatchIn(1, 20, batch.M_max, 0, batch.S, 1, 1, 1, 1, 1, 1)  # empty left lane set
    bo = cilqr_b200.solver.BatchOut(1, 1, 1, None, None, None, None, None, None, None, 0, None)
    assert L.cilqr_plan_batch(solver._h, C.byref(bi), C.byref(bo)) == -1
    bo2 = cilqr_b200.solver.BatchOut(None, 1, 1, None, None, None, None, None, None, None, 0, None)  # null output
    bi2 = cilqr_b200.solver.BatchIn(1, 20, batch.M_max, batch.S, batch.S, 1, 1, 1, 1, 1, 1)
    assert L.cilqr_plan_batch(solver._h, C.byref(bi2), C.byref(bo2)) == -1
    bi3 = cilqr_b200.solver.BatchIn(1, 100000, batch.M_max, batch.S, batch.S, 1, 1, 1, 1, 1, 1)
    assert L.cilqr_plan_batch(solver._h, C.byref(bi3), C.byref(bo)) == -4
