#!/usr/bin/env python
"""Benchmark of the CILQR hot path (BASELINE.json: "CILQR trajectories/sec (100-step horizon)").

A *step* is one pass of the solver over one batch of synthetic random_pedestrian-style scenarios
(BASELINE.json configs[2]: 65 536 scenarios, horizon N = 100, 20 obstacles each, per GPU; weak
scaling: every rank solves its own 65 536-scenario id range, then one NCCL all-gather of the
result blocks).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W]             product arm (CUDA, C ABI)
  python bench.py --impl reference ...                            reference arm: the CPU restatement
                                                                   of the reference solver (oracle/)
                                                                   on all host threads

value     : converged trajectories / s, inputs resident in HBM, CUDA events around the K steps
e2e       : same metric through the host C ABI (cilqr_plan_batch): pinned host buffers in, H2D +
            solve + D2H inside the timed region (one launch fed by chunked copies behind a watermark)
roofline  : algorithmic HBM bytes of the solve kernel / its CUDA-event time vs MEASURED_PEAKS.json
cpu_baseline: the oracle (a port, bit-identical to the reference's own solver source compiled against an
            Eigen stand-in in oracle/_ref, whose rate is reported beside it) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "converged_cilqr_trajectories_per_sec_N100"
UNIT = "traj/s"
SEED = 20260101 + 2  # seed = 20260101 + config index


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=65536)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--obstacles", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=4096, help="scenarios in the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=2048, help="scenarios per step of --impl reference")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="batches in flight in the timed region (one solver handle + stream each)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-corridor", action="store_true", help="skip the corridor-builder side measurement")
    ap.add_argument("--no-dp", action="store_true", help="skip the DP-planner side measurement")
    ap.add_argument("--corridor-base", type=int, default=2048,
                    help="scenarios generated on the host for the corridor measurement (tiled on the device)")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.batch_per_gpu} random_pedestrian-style scenarios per GPU, horizon N={a.horizon}, "
            f"{a.obstacles} obstacles (M_max=20 half-planes/knot, S=40 lane segments/side), seed {SEED}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_baseline(a, kind_note=""):
    """The oracle (C restatement of the reference solver) on all host threads, bounded sample."""
    from cilqr_b200 import scenarios
    from oracle import binding as oracle
    n = a.cpu_sample
    batch = scenarios.generate(SEED, 0, n, N=a.horizon, n_obs=a.obstacles)
    cores = os.cpu_count() or 1
    oracle.solve_batch(batch.slice(0, min(n, 4 * cores)), nthreads=cores)  # warm the library / caches
    t = time.perf_counter()
    _, _, st, conv = oracle.solve_batch(batch, nthreads=cores)
    dt = time.perf_counter() - t
    res = {"value": conv / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"first {n} scenarios of the bench workload, C restatement of the reference solver "
                     f"(gcc -O2, double), {cores} pthreads, {dt:.2f} s wall; mean iterations {st[:, 1].mean():.2f}"}
    res["compiled_reference"] = compiled_reference_rate(batch, st)
    return res


def compiled_reference_rate(batch, status, n: int = 48):
    """The reference's OWN ilqr_optimizer.cc (oracle/_ref/libcilqr_ref_solver.so: compiled unmodified in the build
    container against an Eigen stand-in; bit-identical to the port) on one thread over a few scenarios -- reported for
    transparency, not used as the baseline: the naive stand-in makes it ~3x slower than the port."""
    try:
        from oracle import ref_binding as rb
        if not os.path.exists(rb.SOLVER_LIB_PATH):
            return None
        n = min(n, batch.B)
        t = time.perf_counter()
        for b in range(n):
            rb.ilqr_solve(batch, b)
        dt = time.perf_counter() - t
        conv = int((status[:n, 0] <= 2).sum())
        return {"value": conv / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": f"first {n} scenarios, the reference's own solver source compiled against an Eigen stand-in, "
                          f"1 thread, {dt:.2f} s"}
    except Exception:  # the prebuilt library is optional
        return None


def corridor_measurement(solver, a, dev, stream, peak):
    """Side measurement (rank 0): the batched Corridor::Plan kernel (the step before the
    solve) at the bench shape, next to its CPU restatement on one host core.  Not part of `value`."""
    import torch
    from cilqr_b200 import scenarios
    from cilqr_b200.solver import default_corridor_config
    from oracle import corridor_binding as cb
    base, N, M = min(a.corridor_base, a.batch_per_gpu), a.horizon, 20
    _, ci = scenarios.generate_with_obstacles(SEED, 0, base, N=N, n_obs=a.obstacles)
    rep = max(1, a.batch_per_gpu // base)
    B, K, P = base * rep, ci.K, ci.P_max
    tile = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev).repeat(rep, *([1] * (x.ndim - 1)))  # noqa: E731
    traj, pts, cnt = tile(ci.traj), tile(ci.obs_points), tile(ci.obs_cnt)
    cor = torch.zeros(B, K, M, 3, dtype=torch.float64, device=dev)
    ccnt = torch.zeros(B, K, dtype=torch.int32, device=dev)
    code = torch.zeros(B, K, dtype=torch.int32, device=dev)
    cfg = default_corridor_config(point_cap=4 * a.obstacles + 16)
    ms = []
    for _ in range(4):
        torch.cuda.synchronize()
        solver.corridor_batch_device(B, K, P, M, traj, pts, cnt, cor, ccnt, code, cfg=cfg, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        ms.append(solver.corridor_last_kernel_ms())
    kms = float(np.mean(ms[1:]))
    # algorithmic bytes: trajectory + valid obstacle points + counts in, planes + counts + codes out
    alg = traj.numel() * 8 + int(cnt.sum().item()) * 16 + cnt.numel() * 4 + int(ccnt.sum().item()) * 24 + ccnt.numel() * 8
    fails = int((code != 0).sum().item())
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("corridor", {}).get("dram_bytes_per_launch")
    n_cpu = min(base, 512)
    t0 = time.perf_counter()
    ocor, ocnt, _, ocode = cb.plan_batch(ci.traj[:n_cpu], ci.obs_points[:n_cpu], ci.obs_cnt[:n_cpu], M)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(ccnt[:n_cpu].cpu().numpy(), ocnt))
    return {"kernel": "corridor_build_kernel", "knots_per_launch": B * K, "kernel_ms": kms,
            "traj_per_s": B / kms * 1e3, "knots_per_s": B * K / kms * 1e3, "failed_knots": fails,
            "planes_per_knot": float(ccnt.double().mean().item()),
            "roofline": {"bound": "hbm", "achieved": alg / kms / 1e6, "peak": peak, "unit": "GB/s",
                         "frac": alg / kms / 1e6 / peak, "algorithmic_bytes_per_launch": alg, "traffic": traffic},
            "cpu_baseline": {"value": n_cpu / cpu_s, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"first {n_cpu} scenarios x {K} knots, oracle/corridor_oracle.c, 1 thread, "
                                       f"{cpu_s:.2f} s; plane counts equal to the GPU's: {same}"},
            "note": f"{base} generated scenarios tiled x{rep} on the device; one thread per knot, point_cap {cfg.point_cap}"}


def dp_measurement(solver, a, dev, stream, peak=None):
    """Side measurement (rank 0): the batched DpPlanner::Plan kernel (the first stage of the
    planner) on random_pedestrian-style scenes, next to its CPU restatement on one host core.  Not part of `value`."""
    import torch
    from cilqr_b200 import scenarios
    from cilqr_b200.solver import dp_num_knots
    from oracle import dp_binding as dpo
    base, rep = 512, 16
    db = scenarios.generate_dp(SEED, base, n_obs=11)
    barrier = dpo.build_barrier(db.ref)
    B, K = base * rep, dp_num_knots()
    one = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
    tile = lambda x: one(x).repeat(rep, *([1] * (x.ndim - 1)))  # noqa: E731
    ref, bar = one(db.ref), one(barrier)
    tin = [tile(x) for x in (db.start, db.static_poly, db.static_nv, db.dyn_time, db.dyn_samples, db.dyn_poly, db.dyn_nv)]
    ok = torch.zeros(B, dtype=torch.int32, device=dev)
    coarse = torch.zeros(B, K, 6, dtype=torch.float64, device=dev)
    ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        solver.dp_plan_batch_device(B, len(db.ref), len(barrier), 4, db.static_poly.shape[1], db.dyn_poly.shape[1],
                                    db.dyn_poly.shape[2], ref, bar, *tin, ok, coarse=coarse, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        ms.append(solver.dp_last_kernel_ms())
    kms = float(np.mean(ms[1:]))
    n_cpu = 6
    # CPU side: the reference's OWN DpPlanner (oracle/_ref/libcilqr_ref_dp.so, compiled from the reference sources in
    # the build container; it travels with the repo) when present, else the C restatement -- bit-identical anyway
    kind, okc, cpu_s = "port", [], 0.0
    try:
        from oracle import ref_binding as rb
        if os.path.exists(rb.DP_LIB_PATH):
            t0 = time.perf_counter()
            okc = [rb.dp_plan(db.ref, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                              db.dyn_poly[b], db.dyn_nv[b], *db.start[b])[0] for b in range(n_cpu)]
            cpu_s = time.perf_counter() - t0
            kind = "reference"
    except Exception:  # the prebuilt library is optional
        kind, okc = "port", []
    if kind == "port":
        t0 = time.perf_counter()
        cfg = dpo.default_config()
        okc = []
        for b in range(n_cpu):
            sc = dpo.Scene(db.ref, barrier, db.static_poly[b], db.static_nv[b], db.dyn_time[b], db.dyn_samples[b],
                           db.dyn_poly[b], db.dyn_nv[b])
            okc.append(dpo.plan(sc, *db.start[b], cfg)[0])
        cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(np.array(okc, bool), ok[:n_cpu].cpu().numpy().astype(bool)))
    impl = ("the reference's own DpPlanner::Plan (oracle/_ref)" if kind == "reference" else "oracle/dp_oracle.c")
    # algorithmic bytes: a scene's start + obstacle polygons (static, and every sample of the dynamic ones) in, the
    # coarse trajectory + flag out; the centre line and the barrier are shared by the batch (L2 resident)
    per_scene = 24 + db.static_poly[0].nbytes + db.dyn_poly[0].nbytes + db.dyn_time[0].nbytes + K * 6 * 8 + 4
    alg = per_scene * B + db.ref.nbytes + barrier.nbytes
    roof = None
    if peak:
        roof = {"bound": "hbm", "achieved": alg / kms / 1e6, "peak": peak, "unit": "GB/s", "frac": alg / kms / 1e6 / peak,
                "algorithmic_bytes_per_launch": int(alg), "traffic": None,
                "note": "compute / latency bound (19 670 transitions x ~10 collision-checked points per scene); the HBM "
                        "fraction is reported as defined, not as the limiter"}
    return {"kernel": "dp_plan_kernel", "scenes_per_launch": B, "kernel_ms": kms, "traj_per_s": B / kms * 1e3,
            "planned_ok_fraction": float(ok.double().mean().item()), "roofline": roof,
            "cpu_baseline": {"value": n_cpu / cpu_s, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": f"first {n_cpu} scenes, {impl}, 1 thread, {cpu_s:.2f} s; "
                                       f"ok flags equal to the GPU's: {same}"},
            "note": f"{base} generated scenes (11 obstacles, 81 knots, 5x7x10 lattice) tiled x{rep} on the device; one CTA "
                    "per scene; compute bound (19 670 transitions x 16 collision-checked path points per scene), "
                    "HBM traffic negligible"}


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port) timed on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cilqr_b200 import scenarios
    from oracle import binding as oracle
    cores = os.cpu_count() or 1
    n = a.ref_sample
    batch = scenarios.generate(SEED, 0, n, N=a.horizon, n_obs=a.obstacles)
    for _ in range(a.warmup):
        oracle.solve_batch(batch.slice(0, min(n, 4 * cores)), nthreads=cores)
    conv_total = 0
    t = time.perf_counter()
    st = None
    for _ in range(a.steps):
        _, _, st, conv = oracle.solve_batch(batch, nthreads=cores)
        conv_total += conv
    dt = time.perf_counter() - t
    v = conv_total / dt
    sample = (f"each step = first {n} scenarios of the bench workload (bounded sample), oracle/cilqr_oracle.c "
              f"(C restatement, bit-identical to the reference's own solver source compiled against an Eigen stand-in -- "
              f"oracle/_ref -- and ~3x faster than it), {cores} pthreads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample_per_step": n},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "compiled_reference": compiled_reference_rate(batch, st) if st is not None else None},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    import cilqr_b200
    from cilqr_b200 import scenarios, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product path has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, N = a.batch_per_gpu, a.horizon
    K = N + 1
    total = B * world
    lo, hi = sharding.shard_range(total, rank, world)
    # ---- this rank's shard of the synthetic workload (ids lo..hi-1), host side pinned
    workers = max(1, (os.cpu_count() or 2) // max(1, world))
    t0 = time.time()
    batch = scenarios.generate(SEED, lo, hi - lo, N=N, n_obs=a.obstacles, workers=min(workers, 16))
    gen_s = time.time() - t0
    host_in = [batch.start, batch.coarse, batch.corridor, batch.corridor_cnt, batch.lane_left, batch.lane_right]
    pinned = [torch.from_numpy(x).pin_memory() for x in host_in]
    dev_in = [t.to(dev, non_blocking=True) for t in pinned]
    F = max(1, a.in_flight)
    blocks = [torch.empty(sharding.block_doubles(B, N), dtype=torch.float64, device=dev) for _ in range(F)]
    views = [sharding.carve_block(blk, B, N) for blk in blocks]
    gathered = [torch.empty((world, blocks[0].numel()), dtype=torch.float64, device=dev) if world > 1 else None
                for _ in range(F)]
    torch.cuda.synchronize()

    # One handle = one solve stream + one context workspace.  F handles keep F batches in flight: the
    # persistent solve kernel releases an SM as soon as that SM's scenarios are done, so the next batch's
    # kernel (another stream) fills the SMs that the drain and the straggler tail of this one leave idle.
    solvers = [cilqr_b200.Solver(device=local, N_max=max(N, 100), M_max=batch.M_max, S_max=batch.S, B_max=B)
               for _ in range(F)]
    solver = solvers[0]
    # dedicated (non-default) torch streams: solve kernels, the events that time them and the NCCL
    # all-gathers are enqueued on them (a NULL stream would make the library use its own stream)
    streams = [torch.cuda.Stream(device=dev) for _ in range(F)]
    comm_stream = torch.cuda.Stream(device=dev)
    stream = streams[0]
    torch.cuda.set_stream(stream)
    assert all(st.cuda_stream != 0 for st in streams)
    states, controls, status = views[0]

    def launch(f):
        """one step on slot f: solve on that slot's stream, then (N > 1) the all-gather of its result block,
        ordered after the solve by an event, on the communication stream (collectives stay in issue order)"""
        st = streams[f]
        solvers[f].plan_batch_device(B, N, batch.M_max, batch.S, batch.S, *dev_in, *views[f], stream=st.cuda_stream)
        if world > 1:
            done = torch.cuda.Event()
            done.record(st)
            comm_stream.wait_event(done)
            with torch.cuda.stream(comm_stream):
                sharding.gather_blocks(blocks[f], total, N, out=gathered[f])

    def join():
        """stream 0 waits for everything enqueued on the other streams"""
        for st in streams[1:] + [comm_stream]:
            ev = torch.cuda.Event()
            ev.record(st)
            streams[0].wait_event(ev)

    for i in range(a.warmup):
        launch(i % F)
    join()
    torch.cuda.synchronize()

    def timed(n_flight):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(streams[0])
        for st in streams[1:] + [comm_stream]:
            st.wait_event(t0)
        for i in range(a.steps):
            launch(i % n_flight)
        join()
        t1.record(streams[0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1)

    # ---- (1) one batch in flight: per-launch kernel time (the roofline's denominator) and the serial rate
    launches0 = sum(sv.kernel_launches() for sv in solvers)
    k_ms = []
    serial_ms = 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for _ in range(a.steps):
        ek0, ek1, eg = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        ek0.record(streams[0])
        launch(0)
        ek1.record(streams[0])
        join()
        eg.record(streams[0])
        torch.cuda.synchronize()
        k_ms.append(ek0.elapsed_time(ek1))
        serial_ms += ek0.elapsed_time(eg)
    lib_kernel_ms = solver.last_kernel_ms()  # the library's own event pair around its last solve launch
    # ---- (2) the timed region: K steps, up to F batches in flight
    clocks = ClockSampler(local)
    clocks.start()
    elapsed_ms = timed(F)
    clk = clocks.stop()
    launches = sum(sv.kernel_launches() for sv in solvers) - launches0 - a.steps
    conv_local = int((status[:, 0] <= 2).sum().item())  # every step solves the same scenarios
    for v in views[1:]:
        assert int((v[2][:, 0] <= 2).sum().item()) == conv_local, "slots disagree on the same scenarios"
    iters_mean = float(status[:, 1].mean().item())
    t = torch.tensor([elapsed_ms, float(np.mean(k_ms)), serial_ms], dtype=torch.float64, device=dev)
    c = torch.tensor([conv_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    elapsed_ms, kmean_ms, serial_ms = float(t[0]), float(t[1]), float(t[2])
    conv_total = int(c[0])
    value = conv_total * a.steps / (elapsed_ms * 1e-3)
    value_serial = conv_total * a.steps / (serial_ms * 1e-3)

    # ---- e2e through the host C ABI: pinned host inputs -> H2D -> solve -> D2H
    e2e = None
    if not a.no_e2e:
        hb = scenarios.ScenarioBatch(batch.N, batch.M_max, batch.S, *[p.numpy() for p in pinned])
        outs = [{"states": torch.empty((B, K, 6), dtype=torch.float64).pin_memory().numpy(),
                 "controls": torch.empty((B, N, 2), dtype=torch.float64).pin_memory().numpy(),
                 "status": torch.empty((B, 8), dtype=torch.float64).pin_memory().numpy()} for _ in range(F)]
        for f in range(F):
            solvers[f].plan_batch(hb, out=outs[f])  # warm-up (allocates the staging buffers)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # F host threads, one handle each, every call blocking (ctypes releases the GIL): the H2D stream of one
        # batch and the drain of another overlap, exactly as two planner threads sharing a GPU would
        per_thread = 2
        n_e2e = per_thread * F

        def worker(f):
            for _ in range(per_thread):
                solvers[f].plan_batch(hb, out=outs[f])

        threads = [threading.Thread(target=worker, args=(f,)) for f in range(F)]
        t0 = time.perf_counter()
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        dt = time.perf_counter() - t0
        conv_e = float((outs[0]["status"][:, 0] <= 2).sum())
        assert all(float((o["status"][:, 0] <= 2).sum()) == conv_e for o in outs)
        te = torch.tensor([dt], dtype=torch.float64, device=dev)
        ce = torch.tensor([conv_e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(ce, op=dist.ReduceOp.SUM)
        e2e = {"value": float(ce[0]) * n_e2e / float(te[0]), "unit": UNIT,
               "h2d_bytes_per_step": int(batch.input_bytes()) * world,
               "d2h_bytes_per_step": int(sum(v.nbytes for v in outs[0].values())) * world,
               "steps": n_e2e, "in_flight": F,
               "api": "cilqr_plan_batch (C ABI, host pointers), one blocking call per step from each of "
                      f"{F} host threads (one handle each): one solve launch fed by chunked H2D copies through a "
                      "device watermark; results written by the kernel straight into the pinned host buffers"}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        corridor = None
        if not a.no_corridor:
            corridor = corridor_measurement(solver, a, dev, stream, peak)
        dpm = None
        if not a.no_dp:
            dpm = dp_measurement(solver, a, dev, stream, peak)
        alg_bytes = scenarios.algorithmic_bytes(N, batch.M_max, batch.S) * B
        achieved = alg_bytes / (kmean_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        warps, smem = solver.occupancy(N, batch.S, batch.S)
        res = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "value_one_in_flight": value_serial,
            "config": {"workload": workload_name(a), "total_scenarios_per_step": total, "in_flight": F,
                       "pipelining": f"{F} batches in flight in the timed region (one handle + stream each; a step is "
                                     "still one solve launch over one batch); value_one_in_flight = same K steps, "
                                     "each waited for before the next is enqueued",
                       "l2_policy": f"inputs ({batch.input_bytes() / 1e9:.2f} GB per GPU) exceed the 126 MB L2; no flush",
                       "converged_fraction": conv_total / total, "mean_iterations": iters_mean,
                       "warps_per_sm": warps, "smem_bytes_per_warp": smem, "scenario_gen_s": round(gen_s, 1)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "cilqr_solve_kernel",
                         "kernel_ms": kmean_ms, "kernel_ms_library_events": lib_kernel_ms,
                         "kernel_ms_note": "launch duration with one batch in flight (CUDA events on the launching "
                                           "stream); with F in flight launches overlap and share the SMs",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "fp64 issue/latency bound, not HBM bound: compulsory traffic is inputs once + "
                                 "outputs once; `traffic` (ncu dram bytes per launch) is larger because the "
                                 "scenario contexts live in L2/HBM between solver phases (DESIGN.md)"},
            "clocks": clk, "gpu_launches": int(launches),
            "kernel_ms_per_step_one_in_flight": k_ms,
            "allgather_ms_per_step": (serial_ms - sum(k_ms)) / a.steps if world > 1 else None,
        }
        if e2e:
            res["e2e"] = e2e
        if corridor:
            res["corridor"] = corridor
        if dpm:
            res["dp_planner"] = dpm
        if not a.no_cpu_baseline:
            res["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for sv in solvers:
        sv.close()


if __name__ == "__main__":
    main()
