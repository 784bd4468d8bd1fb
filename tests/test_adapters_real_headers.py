"""The three drop-in classes (include/cilqr/{ilqr_optimizer,corridor,dp_planner}_b200.h) compiled against the
REFERENCE'S OWN headers -- TrajectoryPoint, DiscretizedTrajectory, Environment, LineSegment2d, Polygon2d,
PlannerConfig / IlqrConfig / CorridorConfig / VehicleParam from /root/reference -- instead of the stand-ins of
tests/adapter/stubs.  Only the third-party headers the reference pulls in (Eigen, ROS, OpenCV) come from
oracle/ref_stubs.  This is the closest thing to "builds inside the reference tree" available without ROS."""
import os
import subprocess

import pytest

import cilqr_b200
from cilqr_b200 import build as cbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
FORCE = [x for h in ("array", "algorithm", "tuple", "unordered_map", "limits", "cmath", "numeric", "memory")
         for x in ("-include", h)]
INC = ["-I", os.path.join(ROOT, "oracle", "ref_stubs"), "-I", REF, "-I", os.path.join(ROOT, "include")]

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference is not present")


@pytest.mark.parametrize("src", ["corridor_callsite.cc", "dp_callsite.cc"])
def test_adapter_compiles_against_the_reference_headers(tmp_path, src):
    out = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Werror=return-type"] + FORCE + INC +
                         ["-c", "-o", str(tmp_path / "o.o"), os.path.join(ROOT, "tests", "adapter", "real", src)],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_ilqr_adapter_demo_builds_and_runs_against_the_reference_types(tmp_path):
    """tests/adapter/adapter_demo.cc (the TrajectoryPlanner call site) built with the reference's real
    DiscretizedTrajectory / LineSegment2d / config structs and linked against the reference's own geometry sources."""
    import torch
    cbuild.build_library()
    exe = str(tmp_path / "adapter_demo_real")
    srcs = [os.path.join(ROOT, "tests", "adapter", "adapter_demo.cc")] + [
        os.path.join(REF, "algorithm", p) for p in ("math/vec2d.cpp", "math/line_segment2d.cpp", "math/math_utils.cpp",
                                                    "utils/discretized_trajectory.cpp")]
    out = subprocess.run(["g++", "-std=c++14", "-O1"] + FORCE + INC + srcs + ["-o", exe, "-L", cbuild.LIB_DIR,
                         "-lcilqr_b200", f"-Wl,-rpath,{cbuild.LIB_DIR}"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    if torch.cuda.is_available():
        pytest.skip("run covered by tests/test_adapter.py on a GPU box")
    # no device here: the call site must get `false` and an empty trajectory, loudly (no CPU fallback)
    import numpy as np
    from cilqr_b200 import scenarios
    batch = scenarios.generate(5, 0, 1, N=30)
    parts = [np.array([batch.N, batch.M_max, batch.lane_left.shape[1], batch.lane_right.shape[1]], dtype=np.float64),
             batch.start[0].ravel(), batch.coarse[0].ravel(), batch.corridor_cnt[0].astype(np.float64).ravel(),
             batch.corridor[0].ravel(), batch.lane_left[0].ravel(), batch.lane_right[0].ravel()]
    np.concatenate(parts).astype(np.float64).tofile(tmp_path / "s.bin")  # the format of tests/adapter/adapter_demo.cc
    r = subprocess.run([exe, str(tmp_path / "s.bin"), str(tmp_path / "r.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "cilqr_create failed" in r.stderr and "no CPU fallback" in r.stderr
