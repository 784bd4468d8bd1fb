"""Data-parallel sharding of a scenario batch over the GPUs of one box.

Scenarios are independent, so rank r of G solves the contiguous id range
``[r*B/G, (r+1)*B/G)`` with no data-path communication; the only collective is one all-gather of the
per-shard result block over NCCL / NVLink (``gloo`` in the CPU tests).  A result block is one flat
allocation ``[states b*K*6 | controls b*N*2 | status b*8]`` (doubles) that the solve kernel writes
in place, so the collective moves it without a packing pass.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first ``total % world`` ranks get one extra scenario."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def block_doubles(b: int, N: int) -> int:
    return b * (6 * (N + 1) + 2 * N + 8)


def carve_block(block: torch.Tensor, b: int, N: int):
    """flat result block -> views (states [b,K,6], controls [b,N,2], status [b,8])"""
    K = N + 1
    o1, o2 = b * K * 6, b * K * 6 + b * N * 2
    return block[:o1].view(b, K, 6), block[o1:o2].view(b, N, 2), block[o2:o2 + b * 8].view(b, 8)


def gather_blocks(local: torch.Tensor, total: int, N: int, group=None, out: torch.Tensor | None = None):
    """All-gathers the per-rank result blocks.  Returns ``(buf [world, max_block], sizes)`` where
    ``sizes[r]`` is the number of scenarios of rank r; ``carve_block(buf[r], sizes[r], N)`` gives
    rank r's results.  Uneven shards are padded to the largest block for the collective."""
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]
    nmax = block_doubles(max(sizes), N)
    if local.numel() != nmax:
        pad = torch.zeros(nmax, dtype=local.dtype, device=local.device)
        pad[:local.numel()] = local
        local = pad
    if out is None:
        out = torch.empty((world, nmax), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous(), group=group)
    return out, sizes


def assemble(buf: torch.Tensor, sizes, N: int):
    """gathered blocks -> (states [total,K,6], controls [total,N,2], status [total,8]) in id order"""
    parts = [carve_block(buf[r], b, N) for r, b in enumerate(sizes)]
    return tuple(torch.cat([p[i] for p in parts], dim=0) for i in range(3))
