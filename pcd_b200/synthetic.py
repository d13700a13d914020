"""Synthetic inputs for benchmarks, generated without any CPU reference code: random curve points
are k_i * G computed on the GPU (pcdgpu_fixed_base_mul), scalars are seeded numpy draws."""
from __future__ import annotations

import numpy as np

from . import lib as L

R4 = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137
Q4 = 475922286169261325753349249653048451545124879242694725395555128576210262817955800483758081
FIELD_P = {L.FIELD_R4: R4, L.FIELD_Q4: Q4}
#: coordinate field of each curve's (base-field) generator encoding
_R = 1 << 320

# Generators (affine, plain integers): MNT4 G1 is arkworks' generator; the others are derived
# deterministically (smallest x, cofactor cleared) -- same values as pcd_b200/csrc/constants.cuh.
_GEN_WORDS = None


def _parse_generators():
    """Read the generator limbs out of csrc/constants.cuh (Montgomery form already)."""
    import os
    import re
    global _GEN_WORDS
    if _GEN_WORDS is not None:
        return _GEN_WORDS
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "constants.cuh")).read()
    out = {}
    for cid, name in ((0, "GenMnt4G1"), (1, "GenMnt4G2"), (2, "GenMnt6G1"), (3, "GenMnt6G2")):
        body = src[src.index("struct %s {" % name):]
        body = body[:body.index("\n};")]
        words = []
        for coord in ("gen_x", "gen_y"):
            m = re.search(coord + r"\(int i\) \{ constexpr u32 v\[\d+\] = \{([^}]*)\}", body)
            words += [int(w.strip().rstrip("u"), 16) for w in m.group(1).split(",")]
        w = np.array(words, dtype=np.uint32)
        out[cid] = w.view(np.uint64).copy()
    _GEN_WORDS = out
    return out


def generator(curve: int) -> np.ndarray:
    """Affine generator of the curve as u64 limbs (Montgomery coordinates)."""
    return _parse_generators()[curve]


def random_limbs(n: int, field: int, seed: int) -> np.ndarray:
    """n uniform field elements as (n, 5) u64 limbs (rejection sampling on 298-bit draws)."""
    p = FIELD_P[field]
    pl = np.array([(p >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.zeros((n, 5), dtype=np.uint64)
    todo = np.arange(n)
    while len(todo):
        d = rng.integers(0, 2 ** 64, size=(len(todo), 5), dtype=np.uint64)
        d[:, 4] &= np.uint64((1 << (298 - 256)) - 1)
        lt = np.zeros(len(todo), dtype=bool)
        eq = np.ones(len(todo), dtype=bool)
        for i in range(4, -1, -1):
            lt |= eq & (d[:, i] < pl[i])
            eq &= d[:, i] == pl[i]
        out[todo[lt]] = d[lt]
        todo = todo[~lt]
    return out


def random_points_dev(ctx: L.Context, curve: int, n: int, seed: int):
    """n random points of the curve's prime-order group in device memory (torch int64 tensor of
    shape (n, limbs)): k_i * G with k_i uniform, computed by the GPU's fixed-base kernel."""
    import torch
    dev = torch.device("cuda", ctx.device)
    k = torch.from_numpy(random_limbs(n, L.SCALAR_FIELD_OF[0 if curve < 2 else 1], seed).view(np.int64)).to(dev)
    pts = torch.empty((n, L.AFFINE_LIMBS[curve]), dtype=torch.int64, device=dev)
    ctx.fixed_base_mul_dev(curve, generator(curve), k.data_ptr(), n, pts.data_ptr())
    ctx.sync()
    return pts
