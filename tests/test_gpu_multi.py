"""Multi-GPU entry points of the C ABI (cilqr_multi_create / cilqr_plan_sharded, include/cilqr_b200.h): contiguous
scenario shards over the GPUs of one box from one host process, one NCCL all-gather of the result blocks.  Runs with
every GPU the box has (1 on the round-end box: the all-gather is then the trivial one; 2+ under `gpurun --gpus N`)."""
import numpy as np
import pytest

from cilqr_b200 import scenarios

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [300, 7])
def test_sharded_solve_equals_the_single_gpu_solve(solver, B):
    import torch
    import cilqr_b200
    G = torch.cuda.device_count()
    batch = scenarios.generate(17, 0, B, N=40)
    ref = solver.plan_batch(batch, result=True)
    ms = cilqr_b200.solver.MultiSolver(G, N_max=batch.N, M_max=batch.M_max, S_max=batch.S, B_max_per_device=B)
    per = ms.shard_size(B)
    assert per == (B + G - 1) // G
    K, N = batch.N + 1, batch.N
    blk = per * (6 * K + 2 * N + 8)
    gathered = [torch.full((G * blk,), float("nan"), dtype=torch.float64, device=f"cuda:{r}") for r in range(G)]
    out = ms.plan_sharded(batch, gathered=gathered, result=True)
    for k in ("states", "controls", "status", "result"):
        assert np.array_equal(out[k], ref[k]), k  # sharding does not change a single bit
    # every GPU holds every shard's block [states | controls | status]; rows beyond B are zero
    for r in range(G):
        g = gathered[r].cpu().numpy().reshape(G, blk)
        for q in range(G):
            b0 = q * per
            nb = max(0, min(per, B - b0))
            st = g[q, :per * K * 6].reshape(per, K, 6)
            ct = g[q, per * K * 6:per * K * 6 + per * N * 2].reshape(per, N, 2)
            ss = g[q, per * K * 6 + per * N * 2:].reshape(per, 8)
            assert np.array_equal(st[:nb], ref["states"][b0:b0 + nb]) and np.array_equal(ct[:nb], ref["controls"][b0:b0 + nb])
            assert np.array_equal(ss[:nb], ref["status"][b0:b0 + nb])
            assert np.all(st[nb:] == 0) and np.all(ss[nb:] == 0)
    # a second call reuses the communicators and buffers
    out2 = ms.plan_sharded(batch)
    assert np.array_equal(out2["states"], ref["states"])
    ms.close()
    print(f"\n[multi] {G} GPU(s), B={B}: shards of {per}, results and gathered blocks bit-identical to the single-GPU solve")


def test_multi_argument_checks(solver):
    import ctypes as C
    import cilqr_b200
    L = cilqr_b200.load_library()
    h = C.c_void_p()
    p = cilqr_b200.default_params()
    assert L.cilqr_multi_create(C.byref(p), 0, None, 10, 8, 8, 16, C.byref(h)) == -1
    assert L.cilqr_multi_create(C.byref(p), 10 ** 6, None, 10, 8, 8, 16, C.byref(h)) == -3  # more GPUs than the box has
    assert L.cilqr_multi_devices(None) == 0
