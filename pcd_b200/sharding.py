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
    """Groth16 `create_proof_with_reduction` spread over the GPUs of one box, for when proofs are
    NOT independent (a PCD chain proves its steps one after the other): every rank keeps points [lo, hi) of each of
    the five query vectors resident (with window tables), computes the xyzz partial sums of the five MSMs over its
    range, ONE all_gather moves the partials (4 x 160 + 320 bytes per rank on MNT4-298) and rank 0 adds them and
    assembles the proof (pcdgpu_groth16_assemble_partials_dev).  The witness map's a / b / c chains run on ranks
    0 / 1 / 2 (witness_map_by_vector) and h is broadcast.  The constant points (delta, query[0], alpha / beta) ride
    in rank 0's slices as extra (point, scalar) pairs, as in pcdgpu_pk_upload.  The proof is bit-identical to the
    single-GPU one for any number of ranks."""

    def __init__(self, ctx, pk, cm, rank: int, world: int, device, precompute: bool = True, group=None):
        import ctypes
        from . import lib as L
        from .synthetic import FIELD_P
        self.ctx, self.rank, self.world, self.dev, self.group = ctx, rank, world, device, group
        self.pairing = pk.pairing
        self.g1, self.g2 = L.G1_OF[pk.pairing], L.G2_OF[pk.pairing]
        self.ni, self.nv = cm.num_instance_variables, cm.num_instance_variables + cm.num_witness_variables
        self.p = FIELD_P[L.SCALAR_FIELD_OF[pk.pairing]]
        self.R = (1 << 320) % self.p
        keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return ctypes.c_void_p(a.ctypes.data)

        args = []
        for (ptr, col, val) in (cm.a, cm.b, cm.c):
            args += [arr(ptr, np.uint32), arr(col, np.uint32), arr(val, np.uint64)]
        h = ctypes.c_void_p()
        ctx._check(ctx.lib.pcdgpu_r1cs_upload(ctx.h, pk.pairing, cm.num_constraints, self.ni, cm.num_witness_variables,
                                              *args, ctypes.byref(h)))
        self.r1cs = h
        self.n = ctx.lib.pcdgpu_r1cs_domain_size(h)
        # the four MSMs over the assignment run on contexts (streams + scratch) of their own, beside the witness map
        self.side = [L.Context(ctx.device) for _ in range(4)]
        for c in [ctx] + self.side:  # five MSMs side by side: the prover's window / occupancy rules, not a lone MSM's
            c._check(c.lib.pcdgpu_set_msm_side_by_side(c.h, 1))
        import torch
        self.comm_stream = torch.cuda.Stream(device=device)
        l1, l2 = L.AFFINE_LIMBS[self.g1], L.AFFINE_LIMBS[self.g2]
        q = lambda a, w: np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, w)
        a_q, b1_q, b2_q = q(pk.a_query, l1), q(pk.b_g1_query, l1), q(pk.b_g2_query, l2)
        l_q, h_q = q(pk.l_query, l1), q(pk.h_query, l1)
        one = lambda a, w: np.ascontiguousarray(a, dtype=np.uint64).reshape(1, w)

        def part(points, extras, curve):
            lo, hi = shard_range(points.shape[0], world, rank)
            pts = points[lo:hi]
            if rank == 0 and extras:
                pts = np.concatenate([pts] + extras)
            return (lo, hi, L.Bases(ctx, curve, pts, precompute))

        self.a = part(a_q[1:], [one(pk.delta_g1, l1), a_q[:1], one(pk.alpha_g1, l1)], self.g1)
        self.b1 = part(b1_q[1:], [one(pk.delta_g1, l1), b1_q[:1], one(pk.beta_g1, l1)], self.g1)
        self.b2 = part(b2_q[1:], [one(pk.delta_g2, l2), b2_q[:1], one(pk.beta_g2, l2)], self.g2)
        self.l = part(l_q, [one(pk.delta_g1, l1)], self.g1)
        self.h = part(h_q, [], self.g1)

    def _mont(self, vals):
        import torch
        b = b"".join(int(v * self.R % self.p).to_bytes(40, "little") for v in vals)
        return torch.from_numpy(np.frombuffer(b, dtype="<i8").copy().reshape(-1, 5)).to(self.dev)

    def _scalars(self, base, first_row: int, part, extras):
        """device pointer of the scalars of this rank's slice: rows [first_row + lo, first_row + hi) of `base`, with
        rank 0's extra scalars appended (a copy only on rank 0)"""
        import torch
        lo, hi, bases = part
        rows = base[first_row + lo:first_row + hi]
        if self.rank == 0 and extras is not None:
            rows = torch.cat([rows, extras]).contiguous()
        return rows, bases.n

    def prove(self, z_dev, r: int, s: int):
        """z_dev: (num_vars, 5) int64 tensor on this rank's GPU (Montgomery limbs); r, s: plain integers.  Returns the
        proof's affine limbs on rank 0, None elsewhere."""
        import ctypes
        import torch
        import torch.distributed as dist
        from . import lib as L
        ctx, lib, vp = self.ctx, self.ctx.lib, ctypes.c_void_p
        multi = self.world > 1

        def vector_fn(which):
            t = torch.empty((self.n, 5), dtype=torch.int64, device=self.dev)
            ctx._check(lib.pcdgpu_qap_vector_dev(ctx.h, self.r1cs, which, vp(z_dev.data_ptr()), vp(t.data_ptr())))
            return t

        def combine_fn(a, b, c):
            ctx._check(lib.pcdgpu_qap_combine_dev(ctx.h, self.r1cs, vp(a.data_ptr()), vp(b.data_ptr()), vp(c.data_ptr())))
            return a

        p = self.p
        ex_a = self._mont([r, 1, 1]) if self.rank == 0 else None
        ex_b = self._mont([s, 1, 1]) if self.rank == 0 else None
        ex_l = self._mont([(-r * s) % p]) if self.rank == 0 else None
        x1, x2 = L.XYZZ_LIMBS[self.g1], L.XYZZ_LIMBS[self.g2]
        pab = torch.zeros((2, x1), dtype=torch.int64, device=self.dev)   # a, b_g1
        phl = torch.zeros((2, x1), dtype=torch.int64, device=self.dev)   # h, l
        p2 = torch.zeros((1, x2), dtype=torch.int64, device=self.dev)    # b_g2
        keep = []

        def launch(c, part, base, first, extras, out):
            rows, n = self._scalars(base, first, part, extras)
            keep.append(rows)
            c._check(c.lib.pcdgpu_msm_bases_dev(c.h, part[2].h, 0, vp(rows.data_ptr()), 1, n, vp(out.data_ptr())))

        def gather(t):
            if not multi:
                return t
            with torch.cuda.stream(self.comm_stream):  # not behind the witness map / h MSM queued on the main stream
                outs = [torch.empty_like(t) for _ in range(self.world)]
                dist.all_gather(outs, t, group=self.group)
                res = torch.stack(outs).contiguous()
            self.comm_stream.synchronize()
            return res

        torch.cuda.current_stream().synchronize()  # z, the extras and the zeroed partials are ready for every stream
        # the G2 MSM first (the longest), then the three G1 MSMs over the assignment, each on its own context ...
        launch(self.side[0], self.b2, z_dev, 1, ex_b, p2[0])
        launch(self.side[1], self.a, z_dev, 1, ex_a, pab[0])
        launch(self.side[2], self.b1, z_dev, 1, ex_b, pab[1])
        launch(self.side[3], self.l, z_dev, self.ni, ex_l, phl[1])
        # ... while this context runs the witness map (its three chains on ranks 0, 1, 2) and then the MSM over h
        h = witness_map_by_vector(vector_fn, combine_fn,
                                  lambda: torch.empty((self.n, 5), dtype=torch.int64, device=self.dev), group=self.group)
        if multi:
            if h is None:
                h = torch.empty((self.n, 5), dtype=torch.int64, device=self.dev)
            dist.broadcast(h, src=0, group=self.group)
        launch(ctx, self.h, h, 0, None, phl[0])
        # first exchange: a, b_g1, b_g2 -- rank 0 starts s g_a + r g1_b (4 ms of one thread) under the h MSM
        for c in self.side[:3]:
            c.sync()
        gab, g2 = gather(pab), gather(p2)
        lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
        rl, sl = lim(r), lim(s)
        asm = self.side[0]  # idle by now; holds the assembly state between begin and finish
        if self.rank == 0:
            asm._check(lib.pcdgpu_groth16_assemble_begin_dev(asm.h, self.pairing, vp(rl.ctypes.data), vp(sl.ctypes.data),
                                                             self.world, vp(gab.data_ptr()), vp(g2.data_ptr())))
        # second exchange: h, l
        self.side[3].sync()
        torch.cuda.current_stream().synchronize()
        ghl = gather(phl)
        if self.rank != 0:
            return None
        out = np.zeros(2 * L.AFFINE_LIMBS[self.g1] + L.AFFINE_LIMBS[self.g2], dtype=np.uint64)
        asm._check(lib.pcdgpu_groth16_assemble_finish_dev(asm.h, self.pairing, self.world, vp(ghl.data_ptr()),
                                                          vp(out.ctypes.data)))
        return out

    def close(self):
        for part in (self.a, self.b1, self.b2, self.l, self.h):
            part[2].close()
        for c in self.side:
            c.close()
        self.side = []
        self.ctx.lib.pcdgpu_set_msm_side_by_side(self.ctx.h, 0)
        if self.r1cs:
            self.ctx.lib.pcdgpu_r1cs_free(self.r1cs)
            self.r1cs = None
