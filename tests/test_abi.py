"""The C-ABI shared library: loads without a GPU, exports every symbol include/*.h declares, mirrors the
reference's parameter defaults, and fails loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

import cilqr_b200
from cilqr_b200 import solver as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "cilqr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cilqr_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = cilqr_b200.load_library()
    names = _declared_functions()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cilqr_b200.h but not exported"
    assert sorted(S.EXPORTS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", cilqr_b200.lib_path()], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", out), n
    assert lib.cilqr_abi_version() == 4


def test_product_library_does_not_link_the_oracle():
    out = subprocess.run(["nm", "-D", cilqr_b200.lib_path()], capture_output=True, text=True).stdout
    assert "cilqr_oracle" not in out
    # no product source may import, include, open or link anything under oracle/ (comments and docstrings may
    # NAME the checker -- e.g. the strict build documents what it is compared with -- code may not use it)
    for lib in (cilqr_b200.lib_path(), cilqr_b200.solver.lib_path("strict")):
        if os.path.exists(lib):
            ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
            assert "oracle" not in ldd
            assert "cilqr_oracle" not in subprocess.run(["nm", "-D", lib], capture_output=True, text=True).stdout
    for root, _, files in os.walk(os.path.join(ROOT, "cilqr_b200")):
        for f in files:
            text = open(os.path.join(root, f), errors="ignore").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                code = re.sub(r'""".*?"""', "", text, flags=re.S)
                code = re.sub(r"#.*", "", code)
                assert "oracle" not in code.lower(), f
            elif text:
                code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
                code = re.sub(r"//.*", "", code)
                assert "oracle" not in code.lower(), f


def test_struct_layout_matches_header():
    # CilqrParams: 27 doubles + 2 int32; CilqrBatchIn: 5 int32 (+pad) + 6 pointers; CilqrBatchOut: 10 pointers + int32 (+pad) + 1 pointer
    assert C.sizeof(S.Params) == 27 * 8 + 8
    assert C.sizeof(S.BatchIn) == 24 + 6 * 8 + 8 + 2 * 8
    assert C.sizeof(S.BatchOut) == 10 * 8 + 8 + 8
    assert C.sizeof(S.DebugOut) == 18 * 8
    # CilqrCorridorConfig: 6 doubles + int32 (+pad); CilqrCorridorIn: 4 int32 + 3 pointers; CilqrCorridorOut: 4 pointers
    assert C.sizeof(S.CorridorConfig) == 6 * 8 + 8
    assert C.sizeof(S.CorridorIn) == 16 + 3 * 8
    assert C.sizeof(S.CorridorOut) == 4 * 8
    # CilqrDpConfig: 14 doubles; CilqrDpIn: 7 int32 (+pad) + 9 pointers; CilqrDpOut: 6 pointers
    assert C.sizeof(S.DpConfig) == 14 * 8
    assert C.sizeof(S.DpIn) == 32 + 9 * 8
    assert C.sizeof(S.DpOut) == 6 * 8


def test_default_params_equal_the_oracles(oracle):
    p, q = cilqr_b200.default_params(), oracle.default_params()
    for name, _ in S.Params._fields_:
        assert getattr(p, name) == getattr(q, name), name


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(cilqr_b200.CilqrError) as e:
        cilqr_b200.Solver()
    assert e.value.code == S.E_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_create_argument_validation():
    lib = cilqr_b200.load_library()
    h = C.c_void_p()
    p = cilqr_b200.default_params()
    assert lib.cilqr_create(None, 0, 100, 20, 40, 1, C.byref(h)) == S.E_INVALID
    assert lib.cilqr_create(C.byref(p), 0, 0, 20, 40, 1, C.byref(h)) == S.E_INVALID
    p.num_of_disc = 4  # the kernel is specialised for planner_config.h:58's five discs
    assert lib.cilqr_create(C.byref(p), 0, 100, 20, 40, 1, C.byref(h)) == S.E_INVALID
    assert lib.cilqr_strerror(S.E_SMEM).decode().startswith("horizon does not fit")
