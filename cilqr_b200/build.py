"""Builds cilqr_b200/lib/libcilqr_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcilqr_b200.so")
# The parity instrument: the same sources with -DCILQR_STRICT=1 -fmad=false (reference-ordered arithmetic on the
# portable libm of csrc/pm_math.h; see the top of csrc/cilqr_kernel.cuh).  Loaded only by tests / bench parity legs
# (cilqr_b200.Solver(variant="strict")); the product is LIB_PATH.
STRICT_LIB_PATH = os.path.join(LIB_DIR, "libcilqr_b200_strict.so")
# translation units: (source, extra flags).  dp_capi.cu restates double-precision decision logic of the
# reference and is compiled without FMA contraction (-fmad=false); the solver TU keeps nvcc's default.
UNITS = [("cilqr_capi.cu", []), ("dp_capi.cu", ["-fmad=false"]), ("tracker_capi.cu", ["-fmad=false"]), ("multi_capi.cu", [])]
DEPS = ["cilqr_capi.cu", "cilqr_kernel.cuh", "cilqr_strict.cuh", "pm_math.h", "corridor_kernel.cuh", "dp_capi.cu",
        "dp_kernel.cuh", "cilqr_internal.h", "multi_capi.cu", "tracker_capi.cu", "tracker_kernel.cuh", os.path.join("..", "..", "include", "cilqr_b200.h")]
NVCC_COMPILE = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                "-Xcompiler", "-fPIC"]
NVCC_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "--cudart", "static", "-ldl", "-lpthread"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def is_stale(path: str = LIB_PATH) -> bool:
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build_library(force: bool = False, verbose: bool = False, strict: bool = False) -> str:
    out = STRICT_LIB_PATH if strict else LIB_PATH
    if not force and not is_stale(out):
        return out
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    for src, extra in UNITS:
        tag = "_strict" if strict else ""
        obj = os.path.join(LIB_DIR, os.path.splitext(src)[0] + tag + ".o")
        if strict:
            extra = [e for e in extra if e != "-fmad=false"] + ["-DCILQR_STRICT=1", "-fmad=false"]
        cmd = [_nvcc()] + NVCC_COMPILE + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        subprocess.check_call(cmd, cwd=CSRC)
        objs.append(obj)
    subprocess.check_call([_nvcc()] + NVCC_LINK + ["-o", out] + objs, cwd=CSRC)
    return out


ADAPTER_DEMO = os.path.join(LIB_DIR, "adapter_demo")


def build_adapter_demo(force: bool = False) -> str:
    """Compiles tests/adapter/adapter_demo.cc -- the reference's C++ call site driven through the
    header-compatible planning::IlqrOptimizer (include/cilqr/ilqr_optimizer_b200.h) -- against the
    type stand-ins in tests/adapter/stubs (Eigen/ROS are not installed here) and links it to the
    product library."""
    root = os.path.dirname(_HERE)
    src = os.path.join(root, "tests", "adapter", "adapter_demo.cc")
    hdr = os.path.join(root, "include", "cilqr", "ilqr_optimizer_b200.h")
    if (not force and os.path.exists(ADAPTER_DEMO)
            and os.path.getmtime(ADAPTER_DEMO) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(LIB_PATH))):
        return ADAPTER_DEMO
    build_library()
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(root, "include"),
           "-I", os.path.join(root, "tests", "adapter", "stubs"), src, "-o", ADAPTER_DEMO,
           "-L", LIB_DIR, "-lcilqr_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return ADAPTER_DEMO


CORRIDOR_DEMO = os.path.join(LIB_DIR, "corridor_demo")


def build_corridor_demo(force: bool = False) -> str:
    """Compiles tests/adapter/corridor_demo.cc -- the reference's Corridor call site driven through the
    header-compatible planning::Corridor (include/cilqr/corridor_b200.h) -- against the stand-ins."""
    root = os.path.dirname(_HERE)
    src = os.path.join(root, "tests", "adapter", "corridor_demo.cc")
    hdr = os.path.join(root, "include", "cilqr", "corridor_b200.h")
    if (not force and os.path.exists(CORRIDOR_DEMO)
            and os.path.getmtime(CORRIDOR_DEMO) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(LIB_PATH))):
        return CORRIDOR_DEMO
    build_library()
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(root, "include"),
           "-I", os.path.join(root, "tests", "adapter", "stubs"), src, "-o", CORRIDOR_DEMO,
           "-L", LIB_DIR, "-lcilqr_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return CORRIDOR_DEMO


DP_DEMO = os.path.join(LIB_DIR, "dp_demo")


def build_dp_demo(force: bool = False) -> str:
    """Compiles tests/adapter/dp_demo.cc -- the reference's DpPlanner call site driven through the
    header-compatible planning::DpPlanner (include/cilqr/dp_planner_b200.h) -- against the stand-ins."""
    root = os.path.dirname(_HERE)
    src = os.path.join(root, "tests", "adapter", "dp_demo.cc")
    hdr = os.path.join(root, "include", "cilqr", "dp_planner_b200.h")
    if (not force and os.path.exists(DP_DEMO)
            and os.path.getmtime(DP_DEMO) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(LIB_PATH))):
        return DP_DEMO
    build_library()
    cmd = ["g++", "-std=c++14", "-O2", "-Wall", "-Wextra", "-I", os.path.join(root, "include"),
           "-I", os.path.join(root, "tests", "adapter", "stubs"), src, "-o", DP_DEMO,
           "-L", LIB_DIR, "-lcilqr_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return DP_DEMO


if __name__ == "__main__":
    import sys
    print(build_library(force=True, verbose=True, strict="strict" in sys.argv[1:]))
