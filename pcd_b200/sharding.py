"""Sharding of the proving path across the GPUs of a box: one MSM by point range, and the witness
map's three independent vectors by GPU.

Point-range sharding of one MSM:

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


# ---- the witness map's three vectors on different GPUs ----------------------------------------------------
def vector_owner(which: int, world: int) -> int:
    """Rank that computes vector `which` (0: A z, 1: B z, 2: C z) of R1CStoQAP::witness_map.  Rank 0 combines, so
    with three or more ranks it keeps a and the others go to ranks 1 and 2; with two ranks rank 1 takes b."""
    if world <= 0 or not (0 <= which <= 2):
        raise ValueError("bad vector / world size")
    return which % world if world < 3 else which


def witness_map_by_vector(vector_fn: Callable[[int], "object"], combine_fn: Callable[["object", "object", "object"], "object"],
                          alloc_fn: Callable[[], "object"], group=None):
    """a, b, c chains (iFFT -> coset FFT) on different ranks, results sent to rank 0 for (a b - c) / Z and the coset
    iFFT.  vector_fn(which) -> tensor with this rank's result of stage 1; alloc_fn() -> empty tensor of that shape
    (receive buffer on rank 0); combine_fn(a, b, c) -> h.  Two point-to-point transfers of n x 40 bytes; a single
    NTT is never split.  Returns h on rank 0 and None elsewhere."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    owners = [vector_owner(k, world) for k in range(3)]
    mine = {k: vector_fn(k) for k in range(3) if owners[k] == rank}
    if rank == 0:
        vecs = []
        for k in range(3):
            if owners[k] == 0:
                vecs.append(mine[k])
            else:
                buf = alloc_fn()
                dist.recv(buf, src=owners[k], group=group)
                vecs.append(buf)
        return combine_fn(*vecs)
    for k, t in mine.items():
        dist.send(t, dst=0, group=group)
    return None


# ---- one Groth16 proof over several GPUs ------------------------------------------------------------------
class ShardedGroth16:
    """Groth16 `create_proof_with_reduction` spread over the GPUs of one box, for when proofs are NOT
    independent (a PCD chain proves its steps one after the other).  Everything happens inside libpcdgpu.so
    (pcdgpu_comm_init / pcdgpu_pk_upload_sharded / pcdgpu_groth16_prove_sharded): every rank keeps points [lo, hi)
    of each of the five query vectors resident (with window tables), computes the xyzz partial sums of the five MSMs
    over its range, two NCCL all-gathers of 160 - 480 byte points move them, and every rank assembles the proof.  No
    host bounce, no torch collective on the data path: torch.distributed is only used (once, at construction) to
    broadcast the 128-byte NCCL id.  The proof is bit-identical to the single-GPU one for any number of ranks."""

    def __init__(self, ctx, pk, cm, rank: int, world: int, device=None, precompute: bool = True, group=None):
        from . import snark
        self.ctx, self.rank, self.world = ctx, rank, world
        r, w = ctx.comm_info()
        if w == 1 and world > 1:
            ctx.comm_init_torch(group)
        elif (r, w) != (rank, world):
            raise ValueError("the context's communicator is rank %d of %d, not %d of %d" % (r, w, rank, world))
        self.g = snark.Groth16(ctx, pk.pairing)
        self.index = self.g.index(pk, cm, precompute=precompute, sharded=True)

    def prove(self, z_dev, r: int, s: int):
        """z_dev: torch tensor (or anything with data_ptr()) holding the assignment on this rank's GPU; r, s: Python
        integers.  Collective.  Returns the proof's affine limbs (A || B || C) on every rank."""
        lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
        return self.g.create_proof_sharded_dev(self.index, z_dev.data_ptr(), lim(r), lim(s)).affine_limbs()

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None
