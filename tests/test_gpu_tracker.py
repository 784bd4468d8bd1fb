"""GPU parity of the batched Tracker initial guess (cilqr_tracker_batch / _device) against oracle/tracker_oracle.c --
itself bit-identical to the reference's own tracker.cc (tests/test_reference_pins.py); the kernel's source is
bit-identical to the oracle in host emulation (tests/test_device_code_emulation.py).  On the GPU only CUDA's
cos / sin / tan / hypot / fmod differ from glibc in the last ulp; the tracking controller damps such differences.
Bar: ok flags identical, every field of every trajectory point within 1e-9 relative."""
import os

import numpy as np
import pytest

from cilqr_b200 import scenarios

pytestmark = pytest.mark.gpu


def _coarse_trajectories(seed, B):
    """DP-planned coarse trajectories (the tracker's input, trajectory_planner.cpp:32) from the oracle."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import dp_binding as dpo
    dpo.build()
    db = scenarios.generate_dp(seed, B, n_obs=6)
    barrier = dpo.build_barrier(db.ref)

    def one(b):
        sc = dpo.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b], db.dyn_poly[b],
                       db.dyn_nv[b])
        return dpo.plan(sc, *db.start[b])

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
        res = list(ex.map(one, range(B)))
    keep = [b for b, r in enumerate(res) if r[0] and not np.isnan(r[1]).any()]
    coarse = np.ascontiguousarray([res[b][1] for b in keep])
    rng = np.random.default_rng(seed)
    start4 = np.stack([db.start[keep, 0] + rng.normal(0, 0.2, len(keep)), db.start[keep, 1] + rng.normal(0, 0.2, len(keep)),
                       db.start[keep, 2] + rng.normal(0, 0.03, len(keep)), rng.uniform(3.0, 12.0, len(keep))], axis=1)
    return coarse, np.ascontiguousarray(start4)


def test_tracker_parity_with_oracle(solver):
    from oracle import tracker_binding as tb
    coarse, start4 = _coarse_trajectories(11, 96)
    B, K = coarse.shape[:2]
    got = solver.tracker_batch(start4, coarse)
    ref = [tb.plan(tb.start_record(start4[b]), coarse[b]) for b in range(B)]
    oks = np.array([r[0] for r in ref])
    assert np.array_equal(got["ok"].astype(bool), oks) and oks.all()
    T = np.stack([r[1] for r in ref])
    err = np.abs(got["trajectory"] - T) / (np.abs(T) + 1.0)
    bit = sum(np.array_equal(got["trajectory"][b], T[b]) for b in range(B))
    print(f"\n[tracker parity] B={B} K={K}: ok {int(oks.sum())}/{B}; max rel err {err.max():.2e}; bit-identical trajectories "
          f"{bit}/{B}; kernel {solver.tracker_last_kernel_ms():.1f} ms; DARE iterations per plan ~{ref[0][2]}")
    assert err.max() < 1e-9
    for b in range(0, B, 7):  # the InitGuess copy (ilqr_optimizer.cc:107-139)
        X, U = tb.init_guess(got["trajectory"][b])
        assert np.array_equal(got["guess_states"][b], X) and np.array_equal(got["guess_controls"][b], U)
    # (tracking quality is the reference's: with start speeds of 3-12 m/s the simulated vehicle ends up to ~20 m from the
    # coarse end point -- "Too hard to tun parmes", ilqr_optimizer.cc:106 -- identically in the reference, oracle and kernel)


def test_tracker_feeds_the_solver_on_the_device(solver, oracle):
    """DpPlanner result -> Tracker -> IlqrOptimizer with init_mode = CILQR_INIT_GUESS (the commented-out line
    ilqr_optimizer.cc:168) on the device, against the oracle chain tracker_oracle -> cilqr_oracle(init_mode 2)."""
    import torch
    import cilqr_b200
    from oracle import tracker_binding as tb
    dev = torch.device("cuda:0")
    coarse, start4 = _coarse_trajectories(12, 48)
    B, K = coarse.shape[:2]
    N = K - 1
    # corridor / lanes for the solve: the synthetic generator's constraints are tied to its own coarse path, so build
    # simple ones here: the +-10 m box around every coarse point (AddCorridorPoints' box, corridor.cc:89-120) and the
    # road's own lane boundaries
    th = coarse[:, :, 4]
    c, s = np.cos(th), np.sin(th)
    px, py = coarse[:, :, 2], coarse[:, :, 3]
    M = 8
    cor = np.zeros((B, K, M, 3))
    for f, (bx, by) in enumerate(((c, s), (-s, c), (-c, -s), (s, -c))):
        cor[:, :, f, 0], cor[:, :, f, 1] = 20.0 * bx, 20.0 * by
        cor[:, :, f, 2] = 20.0 * (bx * px + by * py + 10.0)
    cnt = np.full((B, K), 4, np.int32)
    rd = scenarios.road("gentle")
    lanes = [np.ascontiguousarray(np.broadcast_to(scenarios._lane_constraints(rd.lane_pts[side][None], left=(side == 0)),
                                                  (B, len(rd.lane_pts[side]) - 1, 7))) for side in (0, 1)]
    coarse6 = np.ascontiguousarray(coarse[:, :, [2, 3, 4, 6, 7, 9]])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    gx = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    gu = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    ok = torch.zeros(B, dtype=torch.int32, device=dev)
    solver.tracker_batch_device(B, K, t(start4), t(coarse), ok, guess_states=gx, guess_controls=gu)
    solver.synchronize()
    big = cilqr_b200.Solver(device=0, N_max=N, M_max=M, S_max=max(l.shape[1] for l in lanes), B_max=B)
    L = big._L
    bi = cilqr_b200.solver.BatchIn(B, N, M, lanes[0].shape[1], lanes[1].shape[1], *[x.data_ptr() for x in (
        t(start4), t(coarse6), t(cor), t(cnt), t(lanes[0]), t(lanes[1]))], 2, gx.data_ptr(), gu.data_ptr())
    X = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    U = torch.zeros(B, N, 2, dtype=torch.float64, device=dev)
    S = torch.zeros(B, 8, dtype=torch.float64, device=dev)
    import ctypes as C
    bo = cilqr_b200.solver.BatchOut(X.data_ptr(), U.data_ptr(), S.data_ptr(), None, None, None, None, None, None, None, 0, None)
    keep = [t(start4), t(coarse6), t(cor), t(cnt), t(lanes[0]), t(lanes[1])]  # noqa: F841 (the pointers above must stay alive)
    assert ok.all()
    # (BatchIn was built from temporaries: rebuild it from the kept tensors)
    bi = cilqr_b200.solver.BatchIn(B, N, M, lanes[0].shape[1], lanes[1].shape[1], *[x.data_ptr() for x in keep], 2,
                                   gx.data_ptr(), gu.data_ptr())
    assert L.cilqr_plan_batch_device(big._h, C.byref(bi), C.byref(bo), None) == 0
    big.synchronize()
    Xg, Ug, Sg = X.cpu().numpy(), U.cpu().numpy(), S.cpu().numpy()
    big.close()
    # oracle chain
    GX, GU = [], []
    for b in range(B):
        okb, tr, _ = tb.plan(tb.start_record(start4[b]), coarse[b])
        assert okb
        x0, u0 = tb.init_guess(tr)
        GX.append(x0)
        GU.append(u0)
    batch = scenarios.ScenarioBatch(N, M, lanes[0].shape[1], start4, coarse6, cor, cnt, lanes[0], lanes[1])
    Xo, Uo, So, _ = oracle.solve_batch(batch, nthreads=os.cpu_count() or 1, init_mode=2, init_states=np.array(GX),
                                       init_controls=np.array(GU))
    same = (Sg[:, 0] == So[:, 0]) & (Sg[:, 1] == So[:, 1]) & (Sg[:, 7] == So[:, 7])
    e = np.maximum((np.abs(Xg - Xo) / (np.abs(Xo) + 1)).reshape(B, -1).max(axis=1),
                   (np.abs(Ug - Uo) / (np.abs(Uo) + 1)).reshape(B, -1).max(axis=1))
    print(f"\n[tracker->solve] B={B}: identical decision path {same.sum()}/{B}; within 1e-4 on those {(e[same] < 1e-4).mean():.4f}; "
          f"exits {np.bincount(So[:, 0].astype(int), minlength=5).tolist()}, mean iterations {So[:, 1].mean():.2f}")
    assert same.mean() >= 0.9 and (e[same] < 1e-4).mean() >= 0.9
    assert np.isfinite(Xg).all()
