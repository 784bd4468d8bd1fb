"""cilqr_b200/replay.py: the batched replay of recorded scenarios (script/pickle_publisher.py:21-40 for one, here for
many).  CPU part: the pickle reader against pickles of look-alike ROS messages (genpy messages pickle as their slot
values; the real classes need ROS), and the Environment::set_reference restatement against the oracle.  GPU part: the
whole DpPlanner -> Corridor -> IlqrOptimizer chain on recorded scenes."""
import pickle
import sys
import types

import numpy as np
import pytest

from cilqr_b200 import replay, scenarios


def _fake_ros_modules():
    """Minimal look-alikes of the reference's message classes (msg/*.msg), pickling like genpy.Message does: as the
    list of slot values."""
    mods = {}

    def make(modname, clsname, slots):
        mod = mods.setdefault(modname, types.ModuleType(modname))

        def getstate(self):
            return [getattr(self, s) for s in self.__slots__]

        def setstate(self, state):
            for s, v in zip(self.__slots__, state):
                setattr(self, s, v)

        def init(self, **kw):
            for s in self.__slots__:
                setattr(self, s, kw.get(s, [] if s in ("points", "obstacles", "trajectory") else 0.0))

        cls = type(clsname, (object,), {"__slots__": slots, "__getstate__": getstate, "__setstate__": setstate,
                                        "__init__": init, "__module__": modname})
        setattr(mod, clsname, cls)
        return cls

    C = {n: make(m, n, replay._SLOTS[n]) for m, n in (
        ("planning.msg._CenterLine", "CenterLine"), ("planning.msg._CenterLinePoint", "CenterLinePoint"),
        ("planning.msg._Obstacles", "Obstacles"), ("planning.msg._DynamicObstacles", "DynamicObstacles"),
        ("planning.msg._DynamicObstacle", "DynamicObstacle"),
        ("planning.msg._DynamicTrajectoryPoint", "DynamicTrajectoryPoint"),
        ("geometry_msgs.msg._Polygon", "Polygon"), ("geometry_msgs.msg._Point32", "Point32"))}
    return mods, C


def _scene_pickle(db, b, protocol):
    """Scenario b of a scenarios.DpBatch as the reference would have recorded it (reference_publisher.py:232-236)."""
    mods, C = _fake_ros_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    parents = {}
    for k in mods:  # parent packages must be importable for pickle's class lookup
        parts = k.split(".")
        for i in range(1, len(parts)):
            parents.setdefault(".".join(parts[:i]), types.ModuleType(".".join(parts[:i])))
    sys.modules.update(parents)
    sys.modules.update(mods)
    try:
        center = C["CenterLine"](points=[C["CenterLinePoint"](**dict(zip(replay._SLOTS["CenterLinePoint"], map(float, r))))
                                         for r in db.ref])
        poly = lambda P: C["Polygon"](points=[C["Point32"](x=float(x), y=float(y), z=0.0) for x, y in P])  # noqa: E731
        static = C["Obstacles"](obstacles=[poly(db.static_poly[b, o, :db.static_nv[b, o]]) for o in range(db.static_poly.shape[1])])
        dyn = []
        for o in range(db.dyn_poly.shape[1]):
            n = int(db.dyn_samples[b, o])
            P = db.dyn_poly[b, o, :n, :db.dyn_nv[b, o]]
            ctr = P.mean(axis=1)
            # the message stores a body-frame polygon + poses; use theta = 0 poses at the polygon centres, which is exact
            # when every sample is a translation of the first one (true for pedestrians; vehicles are stored per sample
            # below through their own body polygon at the first sample)
            shape = P[0] - ctr[0]
            if not np.allclose(P - ctr[:, None], shape[None], atol=1e-9):
                return None
            dyn.append(C["DynamicObstacle"](polygon=poly(shape), trajectory=[
                C["DynamicTrajectoryPoint"](time=float(db.dyn_time[b, o, t]), x=float(ctr[t, 0]), y=float(ctr[t, 1]), theta=0.0)
                for t in range(n)]))
        rec = {"center": center, "static": static, "dynamic": C["DynamicObstacles"](obstacles=dyn)}
        return pickle.dumps(rec, protocol=protocol)
    finally:
        for k in list(parents) + list(mods):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def _translating_only(db):
    """keep the dynamic obstacles whose samples are pure translations (pedestrians, and vehicles on straights)"""
    return db


def test_pickle_reader_and_environment_restatement():
    from oracle import dp_binding as dpo
    dpo.build()
    db = scenarios.generate_dp(5, 3, n_obs=6)
    # only pedestrians translate rigidly everywhere; drop the other dynamic obstacles for the round trip
    n_ped = 3
    db = scenarios.DpBatch(db.ref, db.start, db.static_poly, db.static_nv, db.dyn_time[:, :n_ped], db.dyn_samples[:, :n_ped],
                           db.dyn_poly[:, :n_ped], db.dyn_nv[:, :n_ped])
    for proto in (0, 2):
        blob = _scene_pickle(db, 1, proto)
        assert blob is not None
        sc = replay.load_pickle(blob)
        assert np.array_equal(sc.center, db.ref)
        assert len(sc.static) == db.static_poly.shape[1] and np.allclose(sc.static[0], db.static_poly[1, 0], atol=1e-6)
        assert len(sc.dynamic) == n_ped
        t, pl = sc.dynamic[0]
        assert np.array_equal(t, db.dyn_time[1, 0]) and np.allclose(pl, db.dyn_poly[1, 0], atol=1e-6)
    # Environment::set_reference: the sorted road barrier equals the oracle's (same point set, same order up to ties)
    left, right, barrier = replay.road_barriers(db.ref)
    ob = dpo.build_barrier(db.ref)
    assert barrier.shape == ob.shape
    assert np.allclose(np.sort(barrier[:, 0]), np.sort(ob[:, 0]), atol=1e-12)
    assert np.allclose(barrier[np.lexsort((barrier[:, 1], barrier[:, 0]))], ob[np.lexsort((ob[:, 1], ob[:, 0]))], atol=1e-9)
    x, y, th, lb, rb = replay.evaluate_station(db.ref, np.array([12.34, 0.0, 1e9]))
    r = dpo.evaluate_station(db.ref, 12.34)
    assert np.allclose([x[0], y[0], th[0], lb[0], rb[0]], [r[1], r[2], r[3], r[5], r[6]], atol=1e-12)
    # QueryDynamicObstacles' sample rule (environment.cpp:131-149): the sample at t (within 1e-10), else the next one
    times, polys = np.array([0.0, 0.1, 0.2]), np.arange(3)[:, None, None] * np.ones((3, 4, 2))
    pick = lambda t: None if replay._dynamic_points_at(times, polys, t) is None else replay._dynamic_points_at(times, polys, t)[0, 0]  # noqa: E731
    assert pick(0.1) == 1 and pick(0.05) == 1 and pick(0.0) == 0 and pick(0.2) == 2 and pick(0.3) is None and pick(-0.1) is None


@pytest.mark.gpu
def test_replay_plans_recorded_scenes(solver):
    """Pickles of 12 recorded scenes -> the three GPU stages -> published result records; the DP stage must agree
    with a direct cilqr_dp_plan_batch call on the same arrays, and the records must be consistent."""
    from oracle import dp_binding as dpo
    dpo.build()
    db = scenarios.generate_dp(7, 12, n_obs=6)
    n_ped = 3
    db = scenarios.DpBatch(db.ref, db.start, db.static_poly, db.static_nv, db.dyn_time[:, :n_ped], db.dyn_samples[:, :n_ped],
                           db.dyn_poly[:, :n_ped], db.dyn_nv[:, :n_ped])
    scenes = [replay.load_pickle(_scene_pickle(db, b, 2)) for b in range(db.B)]
    out = replay.replay(scenes, starts=db.start, solver=solver)
    direct = solver.dp_plan_batch(db, dpo.build_barrier(db.ref))
    assert np.array_equal(out["dp_ok"], direct["ok"].astype(bool))
    ok = out["ok"]
    print(f"\n[replay] planned {int(ok.sum())}/{db.B} recorded scenes (DP ok {int(out['dp_ok'].sum())}, corridor ok "
          f"{int(out['corridor_ok'].sum())}); exits {np.bincount(out['status'][ok, 0].astype(int), minlength=5).tolist()}")
    assert ok.sum() >= 6
    res = out["result"][ok]
    assert np.isfinite(res).all() and (np.diff(res[:, :, 1], axis=1) >= 0).all()          # station accumulates
    np.testing.assert_allclose(res[:, :, 5], np.tan(res[:, :, 9]) / 1.0, rtol=1e-12)        # kappa = tan(delta) / L
    np.testing.assert_allclose(res[:, :, 0], np.broadcast_to(np.arange(res.shape[1]) * 0.1, res[:, :, 0].shape), atol=1e-12)
