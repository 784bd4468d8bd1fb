"""GPU tests of the host path's failure behaviour (cilqr_plan_batch, include/cilqr_b200.h).

The reference never returns an un-filled trajectory as a success: `Optimize` writes `*opt_trajectory` on every
exit (ilqr_optimizer.cc:225,238,285,303,319) and its caller treats an empty one as failure
(trajectory_planner.cpp:91-94).  The streaming host path must keep that property when the input transfer
stalls, and it must work when kernel launches are synchronous (ncu, CUDA_LAUNCH_BLOCKING=1, a debugger).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from cilqr_b200 import scenarios

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_starved_input_transfer_is_an_error_not_a_result(solver):
    """Withhold the input watermark beyond scenario 96: the kernel's watchdog must flag the launch, the call must
    return CILQR_E_TIMEOUT, and every scenario that was not solved must carry the NaN sentinel -- while the
    scenarios that did arrive are solved exactly as in a healthy call."""
    import cilqr_b200
    batch = scenarios.generate(31, 0, 300, N=30)
    good = solver.plan_batch(batch)
    s2 = cilqr_b200.Solver(device=0, N_max=batch.N, M_max=batch.M_max, S_max=batch.S, B_max=batch.B)
    s2.debug_host_path(watchdog_ms=200, starve_after=96)
    out = {"states": np.zeros((batch.B, batch.N + 1, 6)), "controls": np.zeros((batch.B, batch.N, 2)),
           "status": np.zeros((batch.B, 8))}
    with pytest.raises(cilqr_b200.CilqrError) as ei:
        s2.plan_batch(batch, out=out)
    assert ei.value.code == cilqr_b200.solver.E_TIMEOUT
    assert "96 of 300" in str(ei.value)
    assert np.isnan(out["status"][96:, 0]).all()
    assert np.array_equal(out["status"][:96], good["status"][:96])
    assert np.array_equal(out["states"][:96], good["states"][:96])
    # the handle recovers: the next healthy call succeeds and is bit-identical
    s2.debug_host_path(watchdog_ms=4000, starve_after=-1)
    again = s2.plan_batch(batch)
    s2.close()
    for k in ("states", "controls", "status"):
        assert np.array_equal(again[k], good[k]), k


def test_device_path_reports_an_incomplete_launch_through_synchronize(solver):
    """Every status row starts as the sentinel; after a healthy device-path launch none is left."""
    import torch
    dev = torch.device("cuda:0")
    batch = scenarios.generate(32, 0, 64, N=20)
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in
         (batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)]
    X = torch.zeros(64, 21, 6, dtype=torch.float64, device=dev)
    U = torch.zeros(64, 20, 2, dtype=torch.float64, device=dev)
    S = torch.zeros(64, 8, dtype=torch.float64, device=dev)
    solver.plan_batch_device(64, 20, batch.M_max, batch.S, batch.S, *t, X, U, S)
    solver.synchronize()  # raises CilqrError(E_TIMEOUT) when scenarios were left unsolved
    assert torch.isfinite(S).all()


_SYNC_SCRIPT = r"""
import sys
sys.path.insert(0, {root!r})
import numpy as np, torch
import cilqr_b200
from cilqr_b200 import scenarios
batch = scenarios.generate(3, 0, 700, N=30)
s = cilqr_b200.Solver(device=0)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
pb = scenarios.ScenarioBatch(batch.N, batch.M_max, batch.S, *[pin(a) for a in (batch.start, batch.coarse,
     batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right)])
o1 = s.plan_batch(pb)            # pinned inputs: asynchronous chunked copies behind the watermark
o2 = s.plan_batch(batch)         # pageable inputs
assert np.isfinite(o1["status"]).all() and (o1["status"][:, 0] <= 4).all()
for k in ("states", "controls", "status"):
    assert np.array_equal(o1[k], o2[k]), k
print("SYNC_OK", float(o1["status"][:, 1].mean()))
"""


def test_host_path_survives_synchronous_launches():
    """CUDA_LAUNCH_BLOCKING=1 makes every launch return only when the kernel has finished -- what a profiler or a
    debugger does.  The host path enqueues all of its input chunks before the launch, so the kernel never waits for
    work the host has not issued yet (round 1 launched first and deadlocked into its watchdog here)."""
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", CILQR_CHUNK="128", CILQR_WATCHDOG_MS="3000")
    r = subprocess.run([sys.executable, "-c", _SYNC_SCRIPT.format(root=ROOT)], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "SYNC_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_one_handle_dp_then_corridor_then_dp_again(solver):
    """on ONE handle the per-frame order DpPlanner -> Corridor (host path) -> DpPlanner used to free
    the planner's scratch and events when the corridor staging buffer grew.  Results of the second DP call must
    equal the first, and closing the handle must not double-free."""
    import cilqr_b200
    from oracle import dp_binding as dp
    dp.build()
    s = cilqr_b200.Solver(device=0)
    db = scenarios.generate_dp(91, 6)
    barrier = dp.build_barrier(db.ref)
    first = s.dp_plan_batch(db, barrier, waypoints=True)
    K = first["trajectory"].shape[1]
    B = db.B
    pts = np.concatenate([np.broadcast_to(db.static_poly[:, None], (B, K) + db.static_poly.shape[1:]).reshape(B, K, -1, 2),
                          np.transpose(db.dyn_poly, (0, 2, 1, 3, 4)).reshape(B, K, -1, 2)], axis=2)
    cnt = np.full((B, K), pts.shape[2], np.int32)
    xyt = np.nan_to_num(first["xytheta"])
    c1 = s.corridor_batch(xyt, pts, cnt, M_max=24)                      # grows the corridor staging buffer
    rd = scenarios.road("gentle")
    s.lane_constraints(rd.lane_pts[0][None, :40], True, 64)
    second = s.dp_plan_batch(db, barrier, waypoints=True)
    c2 = s.corridor_batch(np.concatenate([xyt, xyt]), np.concatenate([pts, pts]), np.concatenate([cnt, cnt]),
                          M_max=24)                                      # grows it again
    third = s.dp_plan_batch(db, barrier, waypoints=True)
    assert s.dp_last_kernel_ms() > 0.0
    s.close()
    for k in first:
        assert np.array_equal(first[k], second[k], equal_nan=True), k
        assert np.array_equal(first[k], third[k], equal_nan=True), k
    assert np.array_equal(c2["corridor"][:B], c1["corridor"]) and np.array_equal(c2["corridor_cnt"][B:], c1["corridor_cnt"])
