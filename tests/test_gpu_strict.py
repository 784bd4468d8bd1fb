"""The STRICT build of the solve kernel against the oracle on the SAME libm: bit for bit.

What this proves.  The production kernel differs from the CPU oracle by rounding only (re-associated sums, fused
multiply-adds, CUDA's libm instead of glibc's), and a CILQR solve amplifies rounding on a few ill-conditioned
scenarios, so production parity has a small tail (tests/test_gpu_parity.py; the oracle shows the same tail against
itself when only its libm or its FMA contraction changes, tests/test_oracle_golden.py).  The strict build
(libcilqr_b200_strict.so: -DCILQR_STRICT=1 -fmad=false) runs the SAME scheduler, contexts, speculative line search,
deferred angle wraps, retirement of blown-up rollouts, help board and output paths, but evaluates every expression
in the reference's order on the portable libm of csrc/pm_math.h -- and must then reproduce
oracle/libcilqr_oracle_pm.so (same restatement, same libm source, pinned arithmetic otherwise) EXACTLY: every
state, control, cost and status word of every scenario, no tolerance.  Any logic slip in the kernel (a wrong
candidate accepted, a stale nearest segment, a mis-ordered exit) would show up here as a bit difference.
"""
import os

import numpy as np
import pytest

from cilqr_b200 import scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def strict():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    import cilqr_b200
    s = cilqr_b200.Solver(device=0, variant="strict")
    yield s
    s.close()


@pytest.fixture(scope="module")
def oracle_pm():
    from oracle import binding_pm
    binding_pm.build()
    return binding_pm


def _bit_report(name, g, o):
    same = np.array_equal(g, o, equal_nan=True)
    if not same:
        bad = np.argwhere(~((g == o) | (np.isnan(g) & np.isnan(o))))
        r = np.abs(g - o) / (np.abs(o) + 1.0)
        print(f"  {name}: {len(bad)} of {g.size} values differ, first at {bad[0].tolist()}, max rel {np.nanmax(r):.2e}")
    return same


def test_strict_stages_are_bit_identical(strict, oracle_pm):
    """Stage dump of the first iteration (cilqr_debug_first_iteration) against the oracle's stages: localises a
    difference to the function that makes it."""
    import torch
    dev = torch.device("cuda:0")
    batch = scenarios.generate(7, 0, 12, N=60)
    B, N, K, S2 = batch.B, batch.N, batch.N + 1, 2 * batch.S
    z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)  # noqa: E731
    dbg = dict(corridor=z(B, K, batch.M_max, 3), lanes=z(B, S2, 3), X0=z(B, K, 6), U0=z(B, N, 2), cost0=z(B, 5),
               A11=z(B, N, 12), Jx=z(B, K, 6), Ju=z(B, N, 2), Hx=z(B, K, 9), Hu=z(B, N, 2), Kg=z(B, N, 12),
               kg=z(B, N, 2), dV=z(B, 2), Xn=z(B, K, 6), Un=z(B, N, 2), costn=z(B, 5),
               nearest=torch.zeros(B, K, 5, 2, dtype=torch.int32, device=dev), gnorm=z(B))
    tin = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
           (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)]
    strict.debug_first_iteration(B, N, batch.M_max, batch.S, batch.S, *tin, dbg)
    torch.cuda.synchronize()
    g = {k: v.cpu().numpy() for k, v in dbg.items()}
    ok = True
    for b in range(B):
        c = oracle_pm.Ctx(batch, b)
        cor, ll, lr = c.constraints()
        X0, U0 = c.iqr()
        lin = c.linearize(X0, U0)
        Ks, ks, dV = c.backward(1.0)
        Xn, Un = c.forward(1.0, X0, U0)
        A, Bm, Hx = lin["A"], lin["B"], lin["Hx"]
        A11 = np.stack([A[:, 0, 2], A[:, 0, 3], A[:, 0, 4], A[:, 0, 5], A[:, 1, 2], A[:, 1, 3], A[:, 1, 4], A[:, 1, 5],
                        A[:, 2, 3], A[:, 2, 4], A[:, 2, 5], Bm[:, 2, 1]], axis=1)
        Hx9 = np.stack([Hx[:, 0, 0], Hx[:, 0, 1], Hx[:, 0, 2], Hx[:, 1, 1], Hx[:, 1, 2], Hx[:, 2, 2], Hx[:, 3, 3],
                        Hx[:, 4, 4], Hx[:, 5, 5]], axis=1)
        Hu2 = np.stack([lin["Hu"][:, 0, 0], lin["Hu"][:, 1, 1]], axis=1)
        mask = np.arange(batch.M_max)[None, :] < batch.corridor_cnt[b][:, None]
        checks = [("corridor", g["corridor"][b][mask], cor[mask]), ("lanes", g["lanes"][b], np.concatenate([ll, lr])),
                  ("X0", g["X0"][b], X0), ("U0", g["U0"][b], U0), ("cost0", g["cost0"][b], c.total_cost(X0, U0)),
                  ("A11", g["A11"][b], A11), ("Jx", g["Jx"][b], lin["Jx"]), ("Ju", g["Ju"][b], lin["Ju"]),
                  ("Hx", g["Hx"][b], Hx9), ("Hu", g["Hu"][b], Hu2), ("Kg", g["Kg"][b], Ks.reshape(N, 12)),
                  ("kg", g["kg"][b], ks), ("dV", g["dV"][b], dV), ("Xn", g["Xn"][b], Xn), ("Un", g["Un"][b], Un),
                  ("costn", g["costn"][b], c.total_cost(Xn, Un))]
        for name, a_, b_ in checks:
            if not _bit_report(f"scenario {b} {name}", a_, b_):
                ok = False
        c.close()
    assert ok, "strict stage dump differs from the oracle (see the report above)"


@pytest.mark.parametrize("N,B,seed,road", [(30, 512, 11, "gentle"), (50, 1024, 20260102, "gentle"), (100, 2048, 20260103, "gentle"),
                                           (200, 96, 13, "gentle"), (80, 256, 20260101, "shipped")])
def test_strict_full_solves_are_bit_identical(strict, oracle_pm, N, B, seed, road):
    """Whole solves: states, controls and all eight status words (exit flag, iterations, five costs, line-search
    hash) of every scenario equal the oracle's bit for bit -- including configs[1] (1 024 x N=50), a 2 048 slice of
    configs[2] and the tight shipped road with its lambda-overflow exits."""
    kw = dict(n_obs=11) if road == "shipped" else {}
    batch = scenarios.generate(seed, 0, B, N=N, road_name=road, **kw)
    out = strict.plan_batch(batch)
    Xo, Uo, So, _ = oracle_pm.solve_batch(batch, nthreads=os.cpu_count() or 1)
    sx = np.array([np.array_equal(out["states"][b], Xo[b], equal_nan=True) for b in range(B)])
    su = np.array([np.array_equal(out["controls"][b], Uo[b], equal_nan=True) for b in range(B)])
    ss = np.array([np.array_equal(out["status"][b], So[b], equal_nan=True) for b in range(B)])
    allsame = sx & su & ss
    print(f"\n[strict] B={B} N={N} {road}: bit-identical scenarios {int(allsame.sum())}/{B} (states {int(sx.sum())}, controls "
          f"{int(su.sum())}, status {int(ss.sum())}); exits {np.bincount(So[:, 0].astype(int), minlength=5).tolist()}, "
          f"mean iterations {So[:, 1].mean():.2f}; first differing: {np.where(~allsame)[0][:8].tolist()}")
    assert allsame.all()


def test_strict_gradient_exit_and_tolerance_free_solves(oracle_pm):
    """Up to 200 iterations with the cost tolerances off (max-iteration, lambda-overflow and gradient-norm exits):
    still bit-identical."""
    import cilqr_b200
    p = cilqr_b200.solver.default_params()
    p.abs_cost_tol = p.rel_cost_tol = 0.0
    po = oracle_pm.default_params()
    po.abs_cost_tol = po.rel_cost_tol = 0.0
    s = cilqr_b200.Solver(params=p, device=0, variant="strict")
    batch = scenarios.generate(5, 0, 192, N=30)
    out = s.plan_batch(batch)
    s.close()
    Xo, Uo, So, _ = oracle_pm.solve_batch(batch, params=po, nthreads=os.cpu_count() or 1)
    same = np.array([np.array_equal(out["states"][b], Xo[b], equal_nan=True) and np.array_equal(out["status"][b], So[b], equal_nan=True)
                     for b in range(batch.B)])
    print(f"\n[strict, tolerances 0] bit-identical {int(same.sum())}/{batch.B}; exits {np.bincount(So[:, 0].astype(int), minlength=5).tolist()}, "
          f"max iterations {int(So[:, 1].max())}")
    assert same.all()


def test_strict_initial_guess_modes_are_bit_identical(strict, oracle_pm):
    """init_mode (CilqrBatchIn): a caller-provided initial guess instead of iqr -- open-loop rollout of given controls
    (OpenLoopRollout, slover/ilqr.h:362-370) and (states, controls) as given (InitGuess, ilqr_optimizer.cc:107-139,
    the commented-out alternative on line :168).  Same bar: bit for bit."""
    batch = scenarios.generate(23, 0, 128, N=50)
    B, N = batch.B, batch.N
    rng = np.random.default_rng(0)
    Xg = batch.coarse.copy()           # what a tracker would hand over: states near the coarse trajectory ...
    Xg[:, 0, :4] = batch.start
    Xg[:, 0, 4:] = 0.0
    Ug = rng.normal(0.0, 0.05, size=(B, N, 2))  # ... and small controls
    for mode, kw in ((1, dict(init_controls=Ug)), (2, dict(init_states=Xg, init_controls=Ug))):
        out = strict.plan_batch(batch, init_mode=mode, init_guess=True, **kw)
        Xo, Uo, So, _ = oracle_pm.solve_batch(batch, nthreads=os.cpu_count() or 1, init_mode=mode, **kw)
        same = np.array([np.array_equal(out["states"][b], Xo[b], equal_nan=True) and np.array_equal(out["controls"][b], Uo[b], equal_nan=True)
                         and np.array_equal(out["status"][b], So[b], equal_nan=True) for b in range(B)])
        print(f"\n[strict, init_mode {mode}] bit-identical {int(same.sum())}/{B}; exits {np.bincount(So[:, 0].astype(int), minlength=5).tolist()}, "
              f"mean iterations {So[:, 1].mean():.2f}")
        assert same.all()
        if mode == 2:  # iter_trajs[0] is the caller's guess
            assert np.array_equal(out["init_states"], Xg) and np.array_equal(out["init_controls"], Ug)


def test_strict_16k_slice_of_the_bench_workload_is_bit_identical(strict, oracle_pm):
    """16 384 scenarios of BASELINE configs[2] (the workload bench.py times, seed 20260103): the size at which the
    persistent scheduler really runs full (128 contexts per CTA, type epochs, hot contexts, help board, drain) --
    still every bit of every scenario.  (tools/strict_full.py does all 65 536: profiles/r02_strict_full_65536.log.)"""
    batch = scenarios.generate(20260103, 0, 16384, N=100, workers=16)
    out = strict.plan_batch(batch)
    Xo, Uo, So, _ = oracle_pm.solve_batch(batch, nthreads=os.cpu_count() or 1)
    same = np.array([np.array_equal(out["states"][b], Xo[b], equal_nan=True) and np.array_equal(out["controls"][b], Uo[b], equal_nan=True)
                     and np.array_equal(out["status"][b], So[b], equal_nan=True) for b in range(batch.B)])
    print(f"\n[strict, bench workload] bit-identical {int(same.sum())}/{batch.B}; strict kernel {strict.last_kernel_ms():.0f} ms; "
          f"exits {np.bincount(So[:, 0].astype(int), minlength=5).tolist()}, max iterations {int(So[:, 1].max())}")
    assert same.all()
