"""The DP planner oracle (oracle/dp_oracle.c) against an independent pure-Python restatement written from the
reference source (oracle/dp_python.py): same scenes, results bit for bit (both are IEEE double with glibc's libm
and follow the reference's operation order)."""
import numpy as np

from cilqr_b200 import scenarios
from oracle import dp_binding as dp
from oracle import dp_python as pyr


def _python_plan(db, barrier, b):
    ref = pyr.Reference(db.ref)
    statics = [db.static_poly[b, o, :db.static_nv[b, o]] for o in range(db.static_poly.shape[1])]
    dynamics = [[(db.dyn_time[b, o, t], db.dyn_poly[b, o, t, :db.dyn_nv[b, o]]) for t in range(db.dyn_samples[b, o])]
                for o in range(db.dyn_poly.shape[1])]
    env = pyr.Environment(pyr.DEFAULT_CFG, barrier, statics, dynamics)
    return pyr.DpPlanner(pyr.DEFAULT_CFG, ref, env).plan(*db.start[b])


def test_c_oracle_equals_python_restatement():
    db = scenarios.generate_dp(123, 1, n_obs=5)
    barrier = dp.build_barrier(db.ref)
    for b in range(db.B):
        sc = dp.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                      db.dyn_poly[b], db.dyn_nv[b])
        ok, traj, cost, wp = dp.plan(sc, *db.start[b])
        ok2, rows, cost2, wp2 = _python_plan(db, barrier, b)
        assert ok == ok2 and cost == cost2
        assert [tuple(w) for w in wp.tolist()] == [tuple(float(v) for v in w) for w in wp2]
        assert np.array_equal(traj[:, :11], np.array(rows), equal_nan=True)


def test_projection_and_station_queries_agree():
    db = scenarios.generate_dp(5, 1)
    ref = pyr.Reference(db.ref)
    rng = np.random.default_rng(0)
    for s in rng.uniform(-3.0, db.ref[-1, 0] + 3.0, size=200):
        assert tuple(dp.evaluate_station(db.ref, s)) == ref.evaluate_station(float(s))
    for _ in range(20):
        i = int(rng.integers(0, len(db.ref)))
        x, y = db.ref[i, 1] + rng.normal(0, 2.0), db.ref[i, 2] + rng.normal(0, 2.0)
        assert tuple(dp.get_projection(db.ref, x, y)) == ref.get_projection(float(x), float(y))
