"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard a batch, each 'solves' its id range
(the oracle stands in for the device solver, which needs a GPU) and one all-gather reassembles the
result blocks in id order -- the same code path bench.py runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cilqr_b200 import scenarios, sharding


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8, 65536, 1048576 + 3):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_block_carving():
    b, N = 5, 7
    blk = torch.arange(sharding.block_doubles(b, N), dtype=torch.float64)
    X, U, S = sharding.carve_block(blk, b, N)
    assert X.shape == (5, 8, 6) and U.shape == (5, 7, 2) and S.shape == (5, 8)
    assert X.data_ptr() == blk.data_ptr() and S[-1, -1] == blk[-1]


def _worker(rank, world, port, total, N, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as oracle
    lo, hi = sharding.shard_range(total, rank, world)
    batch = scenarios.generate(seed, 0, total, N=N, workers=1).slice(lo, hi)
    b = hi - lo
    block = torch.zeros(sharding.block_doubles(b, N), dtype=torch.float64)
    X, U, S = sharding.carve_block(block, b, N)
    Xo, Uo, So, _ = oracle.solve_batch(batch, nthreads=1)
    X.copy_(torch.from_numpy(Xo)); U.copy_(torch.from_numpy(Uo)); S.copy_(torch.from_numpy(So))
    buf, sizes = sharding.gather_blocks(block, total, N)
    Xa, Ua, Sa = sharding.assemble(buf, sizes, N)
    if rank == 0:
        q.put((Xa.numpy(), Ua.numpy(), Sa.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [12, 13])
def test_two_rank_shard_and_allgather(total, oracle):
    N, seed = 16, 31
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, N, seed, q)) for r in range(2)]
    [p.start() for p in procs]
    Xa, Ua, Sa = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    batch = scenarios.generate(seed, 0, total, N=N, workers=1)
    Xo, Uo, So, _ = oracle.solve_batch(batch, nthreads=2)
    assert np.array_equal(Xa, Xo) and np.array_equal(Ua, Uo) and np.array_equal(Sa, So)
