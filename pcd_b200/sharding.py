"""Point-range sharding of one MSM across the GPUs of a box.

sum_i s_i P_i is a sum of independent partial sums: rank g keeps points [lo_g, hi_g) of every query
vector resident, computes an xyzz partial over its range, and the partials (160 / 320 / 480 bytes)
are gathered -- the path's one exchange step -- and added on every rank.  The affine result is
canonical, so it is bit-identical for any number of GPUs.  The collective is a plain
`torch.distributed.all_gather` of a byte tensor (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, n): the first n % world ranks get one extra point."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n, world, r) for r in range(world)]


def gather_partials(partial: np.ndarray, group=None, device=None) -> np.ndarray:
    """all_gather of one rank's xyzz partial(s) (uint64 limbs, any shape) -> array with a leading
    rank axis, in rank order.  Works without an initialised process group (world size 1)."""
    import torch
    import torch.distributed as dist
    p = np.ascontiguousarray(partial, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()):
        return p[None]
    world = dist.get_world_size(group)
    t = torch.from_numpy(p.view(np.int64).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy().view(np.uint64).reshape(p.shape) for o in out])


def sharded_msm(local_partial: Callable[[int, int], np.ndarray], combine: Callable[[np.ndarray], np.ndarray], n: int,
                group=None, device=None) -> np.ndarray:
    """local_partial(lo, hi) -> this rank's xyzz partial over points [lo, hi); combine(parts) -> affine
    sum of the gathered partials (pcd_b200.Context.xyzz_sum on a GPU).  Returns the affine result,
    identical on every rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_range(n, world, rank)
    parts = gather_partials(local_partial(lo, hi), group=group, device=device)
    return combine(parts)
