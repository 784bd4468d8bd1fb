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
    ap.add_argument("--batch-per-gpu", type=int, default=0,
                    help="0 = BASELINE.json as written: 65 536 per GPU (configs[2]); at 8 GPUs 131 072 per GPU = the "
                         "1 048 576-scenario batch of configs[3]")
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--obstacles", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=4096, help="scenarios in the cpu_baseline sample")
    ap.add_argument("--ref-sample", type=int, default=2048, help="scenarios per step of --impl reference")
    ap.add_argument("--in-flight", type=int, default=2,
                    help="batches in flight in the timed region (one solver handle + stream each)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the B=1 latency measurement (configs[0])")
    ap.add_argument("--no-corridor", action="store_true", help="skip the corridor-builder side measurement")
    ap.add_argument("--no-dp", action="store_true", help="skip the DP-planner side measurement")
    ap.add_argument("--corridor-base", type=int, default=2048,
                    help="scenarios generated on the host for the corridor measurement (tiled on the device)")
    return ap.parse_args()


def resolve_batch(a, world: int):
    if a.batch_per_gpu <= 0:
        a.batch_per_gpu = 131072 if world == 8 else 65536
    return a.batch_per_gpu


def workload_name(a):
    cfg = {65536: "configs[2]: ", 131072: "configs[3] (1 048 576 scenarios over 8 GPUs): "}.get(a.batch_per_gpu, "")
    return (f"{cfg}{a.batch_per_gpu} random_pedestrian-style scenarios per GPU, horizon N={a.horizon}, "
            f"{a.obstacles} obstacles (M_max=20 half-planes/knot, S=40 lane segments/side), seed {SEED}")


def kernel_source_hash() -> str:
    """sha1 over the solve kernel's sources: ties profiles/traffic.json to the code it was captured from."""
    import hashlib
    h = hashlib.sha1()
    for f in ("cilqr_kernel.cuh", "cilqr_capi.cu"):
        h.update(open(os.path.join(ROOT, "cilqr_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def fp64_peak_tflops():
    """Non-tensor FP64 peak: measured DFMA rate of tools/microbench/fp64_peak.cu when its result is committed
    (profiles/fp64_peak_b200.json), else 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz."""
    pth = os.path.join(ROOT, "profiles", "fp64_peak_b200.json")
    if os.path.exists(pth):
        return json.load(open(pth))["dfma_tflops"], "measured (profiles/fp64_peak_b200.json, tools/microbench/fp64_peak.cu)"
    return 148 * 64 * 2 * 1.965e9 / 1e12, "nominal: 148 SMs x 64 DFMA/clk/SM x 2 x 1.965 GHz"


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


def cpu_baseline(a, gpu=None):
    """The oracle (C restatement of the reference solver) on all host threads, bounded sample.  With `gpu` =
    (states, controls, status) of the same scenarios from the timed CUDA path, also the parity figures of the
    workload that was timed (-> config.parity)."""
    from cilqr_b200 import scenarios
    from oracle import binding as oracle
    n = a.cpu_sample
    batch = scenarios.generate(SEED, 0, n, N=a.horizon, n_obs=a.obstacles)
    cores = os.cpu_count() or 1
    oracle.solve_batch(batch.slice(0, min(n, 4 * cores)), nthreads=cores)  # warm the library / caches
    t = time.perf_counter()
    Xo, Uo, st, conv = oracle.solve_batch(batch, nthreads=cores)
    dt = time.perf_counter() - t
    parity = None
    if gpu is not None:
        def three(A, Bq):
            (Xa, Ua, Sa), (Xb, Ub, Sb) = A, Bq
            same = (Sa[:, 0] == Sb[:, 0]) & (Sa[:, 1] == Sb[:, 1]) & (Sa[:, 7] == Sb[:, 7])
            rel = lambda x, y: (np.abs(x - y) / (np.abs(y) + 1.0)).reshape(n, -1).max(axis=1)  # noqa: E731
            e = np.maximum(rel(Xa, Xb), rel(Ua, Ub))
            bit = sum(np.array_equal(Xa[b], Xb[b], equal_nan=True) and np.array_equal(Ua[b], Ub[b], equal_nan=True)
                      and np.array_equal(Sa[b], Sb[b], equal_nan=True) for b in range(n))
            return {"identical_path": int(same.sum()), "of": n, "within_1e-4": int((e[same] < 1e-4).sum()),
                    "worst": float(e[same].max()), "median": float(np.median(e[same])),
                    "p99": float(np.quantile(e[same], 0.99)), "bit_identical": int(bit)}
        parity = {"sample": f"first {n} scenarios of the timed workload, GPU (timed path) vs oracle"}
        parity.update(three(gpu, (Xo, Uo, st)))
        parity["same_exit_flag"] = int((gpu[2][:, 0] == st[:, 0]).sum())
        parity["definition"] = ("identical_path = same (exit flag, iteration count, FNV hash of the accepted "
                                "line-search index per iteration); within_1e-4 / worst / median / p99 = max over "
                                "states and controls of |a - b| / (|b| + 1), on identical-path scenarios")
        # Where the tail comes from: the same sample through (1) the strict build of the kernel (reference-ordered
        # arithmetic, portable libm) and (2) the oracle on that same portable libm.  strict == oracle(pm) bit for bit
        # proves the kernel's logic; oracle(pm) vs oracle(glibc) is what exchanging the libm alone does to the
        # REFERENCE algorithm; production vs oracle(glibc) (above) is of that size.
        try:
            import cilqr_b200
            from oracle import binding_pm
            if os.path.exists(cilqr_b200.solver.lib_path("strict")) and os.path.exists(binding_pm._LIB_PATH):
                ss = cilqr_b200.Solver(device=int(os.environ.get("LOCAL_RANK", "0")), N_max=max(a.horizon, 100),
                                       M_max=batch.M_max, S_max=batch.S, B_max=n, variant="strict")
                o = ss.plan_batch(batch)
                ss.close()
                strict = (o["states"], o["controls"], o["status"])
                Xp, Up, Sp, _ = binding_pm.solve_batch(batch, nthreads=cores)
                parity["separation"] = {
                    "strict_gpu_vs_oracle_pm_libm": three(strict, (Xp, Up, Sp)),
                    "oracle_pm_libm_vs_oracle_glibc": three((Xp, Up, Sp), (Xo, Uo, st)),
                    "strict_gpu_vs_oracle_glibc": three(strict, (Xo, Uo, st)),
                    "note": "strict GPU build = libcilqr_b200_strict.so (-DCILQR_STRICT=1 -fmad=false: same scheduler "
                            "and data flow, reference-ordered arithmetic, csrc/pm_math.h); oracle_pm = "
                            "oracle/libcilqr_oracle_pm.so (same restatement on the same pm_math.h)"}
        except Exception as ex:  # the parity instrument is optional for the bench line
            parity["separation"] = {"error": str(ex)[:200]}
    res = {"value": conv / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"first {n} scenarios of the bench workload, C restatement of the reference solver "
                     f"(gcc -O2, double), {cores} pthreads, {dt:.2f} s wall; mean iterations {st[:, 1].mean():.2f}"}
    res["compiled_reference"] = compiled_reference_rate(batch, st)
    return res, parity


def latency_b1(repeat: int = 10, n_scen: int = 16):
    """BASELINE.json configs[0] -- the reference's actual use: ONE ego, one IlqrOptimizer::Plan call per click
    (planning_node.cc:82-88), N = 80, 11 obstacles, the shipped road.  Wall-clock latency of that call through the
    drop-in C++ class (include/cilqr/ilqr_optimizer_b200.h -> cilqr_plan_batch, B = 1, pageable std::vector buffers,
    full iter_trajs / cost() history), next to the CPU port on ONE host core on the same scenarios."""
    import tempfile
    from cilqr_b200 import build as cbuild
    from cilqr_b200 import scenarios
    from oracle import binding as oracle
    exe = cbuild.ADAPTER_DEMO
    if not os.path.exists(exe):
        return None
    N = 80
    batch = scenarios.generate(20260101, 0, n_scen, N=N, n_obs=11, road_name="shipped")
    with tempfile.TemporaryDirectory() as td:
        files = []
        for b in range(n_scen):
            parts = [np.array([N, batch.M_max, batch.lane_left.shape[1], batch.lane_right.shape[1]], dtype=np.float64),
                     batch.start[b].ravel(), batch.coarse[b].ravel(), batch.corridor_cnt[b].astype(np.float64).ravel(),
                     batch.corridor[b].ravel(), batch.lane_left[b].ravel(), batch.lane_right[b].ravel()]
            f = os.path.join(td, f"s{b}.bin")
            np.concatenate(parts).astype(np.float64).tofile(f)
            files.append(f)
        r = subprocess.run([exe, "--latency", str(repeat)] + files, capture_output=True, text=True, timeout=600)
    lat = np.array([[float(x) for x in ln.split()[1:]] for ln in r.stdout.splitlines() if ln.startswith("LAT ")])
    if r.returncode != 0 or len(lat) == 0:
        return {"error": (r.stdout + r.stderr)[-300:]}
    cpu_ms = []
    for b in range(n_scen):
        t0 = time.perf_counter()
        o = oracle.solve(batch, b)
        cpu_ms.append((time.perf_counter() - t0) * 1e3)
    cpu_ms = np.array(cpu_ms)
    g = lat[:, 1]
    return {"config": "configs[0]: single ego, N=80, 11 obstacles, shipped road (emulated by the generator), B=1",
            "api": "planning::IlqrOptimizer::Plan (drop-in class) -> cilqr_plan_batch, B=1",
            "calls": int(len(g)), "scenarios": n_scen, "gpu_ms_p50": float(np.median(g)), "gpu_ms_p99": float(np.quantile(g, 0.99)),
            "gpu_ms_mean": float(g.mean()), "gpu_ms_min": float(g.min()), "mean_iterations": float(lat[:, 3].mean()),
            "cpu_port_ms_p50": float(np.median(cpu_ms)), "cpu_port_ms_p99": float(np.quantile(cpu_ms, 0.99)),
            "cpu_port_ms_mean": float(cpu_ms.mean()), "cpu_cores": 1,
            "gpu_faster": bool(np.median(g) < np.median(cpu_ms))}


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
    traj = torch.zeros(B, K, 13, dtype=torch.float64, device=dev)
    ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        solver.dp_plan_batch_device(B, len(db.ref), len(barrier), 4, db.static_poly.shape[1], db.dyn_poly.shape[1],
                                    db.dyn_poly.shape[2], ref, bar, *tin, ok, coarse=coarse, trajectory=traj,
                                    stream=stream.cuda_stream)
        torch.cuda.synchronize()
        ms.append(solver.dp_last_kernel_ms())
    kms = float(np.mean(ms[1:]))
    # the Tracker initial guess on the planner's own output, chained on the device
    tracker = None
    try:
        from oracle import tracker_binding as tb
        good = torch.nonzero(ok.bool() & torch.isfinite(traj).all(dim=2).all(dim=1)).flatten()
        tg = traj[good].contiguous()
        st4 = torch.cat([tin[0][good], torch.full((len(good), 1), 10.0, dtype=torch.float64, device=dev)], dim=1).contiguous()
        gx = torch.zeros(len(good), K, 6, dtype=torch.float64, device=dev)
        gu = torch.zeros(len(good), K - 1, 2, dtype=torch.float64, device=dev)
        tok = torch.zeros(len(good), dtype=torch.int32, device=dev)
        tms = []
        for _ in range(3):
            torch.cuda.synchronize()
            solver.tracker_batch_device(len(good), K, st4, tg, tok, guess_states=gx, guess_controls=gu, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            tms.append(solver.tracker_last_kernel_ms())
        n_t = 8
        th, sh = tg[:n_t].cpu().numpy(), st4[:n_t].cpu().numpy()
        t0 = time.perf_counter()
        ref_t = [tb.plan(tb.start_record(sh[b]), th[b]) for b in range(n_t)]
        cpu_t = (time.perf_counter() - t0) / n_t
        gx_h = gx[:n_t].cpu().numpy()
        err = max(float(np.abs(gx_h[b] - tb.init_guess(ref_t[b][1])[0]).max()) for b in range(n_t))
        tk = float(np.mean(tms[1:]))
        tracker = {"kernel": "tracker_kernel", "scenes_per_launch": int(len(good)), "kernel_ms": tk,
                   "traj_per_s": len(good) / tk * 1e3, "ok_fraction": float(tok.double().mean().item()),
                   "cpu_baseline": {"value": 1.0 / cpu_t, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"first {n_t} scenes, oracle/tracker_oracle.c (bit-identical to the reference's own "
                                              f"tracker.cc), 1 thread, {cpu_t * 1e3:.1f} ms per plan; max abs difference of the guess "
                                              f"states GPU vs CPU: {err:.1e}"},
                   "note": "800 sequential 10 ms simulation steps per scene, ~50 000 DARE iterations; one thread per scene "
                           "(latency bound by construction); HBM traffic negligible"}
    except Exception as ex:  # side measurement only
        tracker = {"error": str(ex)[:200]}
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
            "planned_ok_fraction": float(ok.double().mean().item()), "roofline": roof, "tracker": tracker,
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
    resolve_batch(a, int(os.environ.get("WORLD_SIZE", "1")))
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
        # NCCL sizes its channel count from the topology it detects; on this pool's boxes it picks few channels and the
        # all-gather of the result blocks runs at ~140 GB/s per GPU.  32 channels: 444 GB/s at 2 GPUs
        # (profiles/r02_k_nccl_channels.txt).  A user's own setting wins.
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "32")
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

    used = set()

    def launch(f):
        """one step on slot f: solve on that slot's stream, then (N > 1) the all-gather of its result block,
        ordered after the solve by an event, on the communication stream (collectives stay in issue order)"""
        used.add(f)
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
    sched = solver.debug_stats()             # scheduler counters of that launch (summed over all warps)
    # ---- (2) the timed region: K steps, up to F batches in flight
    clocks = ClockSampler(local)
    clocks.start()
    elapsed_ms = timed(F)
    clk = clocks.stop()
    launches = sum(sv.kernel_launches() for sv in solvers) - launches0 - a.steps
    conv_local = int((status[:, 0] <= 2).sum().item())  # every step solves the same scenarios
    for f, v in enumerate(views):
        if f in used and f > 0:
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
        per_thread = max(2, -(-a.steps // F))  # about as many steps as the device-resident leg (K = --steps)
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
        # `traffic` and the FP64 operation count come from ONE ncu capture of this kernel at this shape
        # (profiles/traffic.json, written by tools/ncu_traffic.py together with a hash of the kernel sources);
        # a capture of different sources is stale and is reported as null, with the reason
        traffic, flops, cap_note = None, None, "no ncu capture committed"
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("kernel_source_sha1") != kernel_source_hash():
                cap_note = (f"stale: profiles/traffic.json was captured from kernel sources {str(tj.get('kernel_source_sha1'))[:10]} "
                            f"(commit {tj.get('commit')}), this run is {kernel_source_hash()[:10]}")
            elif tj.get("scenarios") != B or tj.get("horizon") != N:
                cap_note = f"capture is of {tj.get('scenarios')} scenarios, N={tj.get('horizon')}: not this shape"
            else:
                traffic = tj.get("dram_bytes_per_launch")
                flops = tj.get("fp64_flops_per_launch")
                cap_note = f"ncu --set full capture of commit {tj.get('commit')} ({tj.get('source')}), same kernel sources"
        fp64_peak, fp64_src = fp64_peak_tflops()
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
                       "warps_per_sm": warps, "smem_bytes_per_warp": smem, "scenario_gen_s": round(gen_s, 1),
                       "scheduler": {**sched, "per_trajectory": {k: round(v / B, 2) for k, v in sched.items()}}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": cap_note,
                         "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                         # secondary ceiling: the kernel is fp64 issue/latency bound, so the distance
                         # from the machine is the FP64 pipe fraction, not the HBM one
                         "flops": {"fp64_flops_per_launch": flops,
                                   "achieved_tflops": (flops / (kmean_ms * 1e-3) / 1e12) if flops else None,
                                   "peak_tflops": fp64_peak, "peak_source": fp64_src,
                                   "fp64_frac": (flops / (kmean_ms * 1e-3) / 1e12 / fp64_peak) if flops else None,
                                   "definition": "2*DFMA + DADD + DMUL thread instructions (ncu smsp__sass_thread_inst_"
                                                 "executed_op_d*_pred_on) of one launch / launch duration"},
                         "peak_source": peak_src, "kernel": "cilqr_solve_kernel",
                         "kernel_ms": kmean_ms, "kernel_ms_library_events": lib_kernel_ms,
                         "kernel_ms_note": "launch duration with one batch in flight (CUDA events on the launching "
                                           "stream); with F in flight launches overlap and share the SMs",
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "fp64 issue/latency bound, not HBM bound: compulsory traffic is inputs once + "
                                 "outputs once; `traffic` (ncu dram bytes per launch) is larger because the "
                                 "scenario contexts live in L2/HBM between solver phases (DESIGN.md)"},
            "clocks": clk, "gpu_launches": int(launches),
            "kernel_ms_per_step_one_in_flight": k_ms,
        }
        if world > 1:
            ag_ms = (serial_ms - sum(k_ms)) / a.steps
            ag_bytes = blocks[0].numel() * 8 * (world - 1)  # received by every GPU per step
            res["config"]["allgather"] = {"ms_per_step": ag_ms, "bytes_received_per_gpu": int(ag_bytes),
                                          "achieved_gbs_per_gpu": ag_bytes / (ag_ms * 1e-3) / 1e9 if ag_ms > 0 else None,
                                          "nvlink_peak_gbs_per_direction": 900.0,
                                          "nccl_min_nchannels": os.environ.get("NCCL_MIN_NCHANNELS"),
                                          "note": "exposed time of the all-gather with one batch in flight (step time "
                                                  "minus solve-kernel time); with batches in flight it overlaps the next solve"}
        if e2e:
            res["e2e"] = e2e
        if corridor:
            res["corridor"] = corridor
        if dpm:
            res["dp_planner"] = dpm
        if not a.no_cpu_baseline:
            n = min(a.cpu_sample, B)
            a.cpu_sample = n
            gpu = (states[:n].cpu().numpy(), controls[:n].cpu().numpy(), status[:n].cpu().numpy())
            res["cpu_baseline"], res["config"]["parity"] = cpu_baseline(a, gpu)
        if not a.no_latency:
            res["latency_b1"] = latency_b1()
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for sv in solvers:
        sv.close()


if __name__ == "__main__":
    main()
