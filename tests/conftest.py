"""Shared fixtures.  Tests marked ``gpu`` need a B200 (run with ``-m gpu``); everything else runs
on CPU (``-m "not gpu"``).  The oracle (oracle/) is the checker; the product path under test is
cilqr_b200/lib/libcilqr_b200.so reached through its C ABI."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")


def _cuda_visible() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (not skip) when they were SELECTED (-m gpu) on a box without a device: a silent
    # skip would read as a pass of a path that never ran.  A plain `pytest` (no -m) on a CPU box skips them
    # with a reason instead of dying in the first one.
    markexpr = (config.getoption("markexpr", "") or "").strip()
    if "gpu" in markexpr or _cuda_visible():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device; run with -m gpu on a B200 (fails loudly there if absent)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def small_batch():
    from cilqr_b200 import scenarios
    return scenarios.generate(7, 0, 12, N=30)


@pytest.fixture(scope="session")
def solver():
    import torch
    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    import cilqr_b200
    s = cilqr_b200.Solver(device=0)
    yield s
    s.close()
