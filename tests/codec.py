"""Byte/limb helpers shared by the tests: hex <-> numpy uint64 limb arrays, golden loaders, and
seeded synthetic inputs.  Test infrastructure."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

R4 = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137
Q4 = 475922286169261325753349249653048451545124879242694725395555128576210262817955800483758081
FIELD_P = {0: R4, 1: Q4}
#: u64 limbs per coordinate-field element / affine point of each curve
COORD_LIMBS = {0: 5, 1: 10, 2: 5, 3: 15}
POINT_LIMBS = {0: 10, 1: 20, 2: 10, 3: 30}
CURVE_ORDER = {0: R4, 1: R4, 2: Q4, 3: Q4}
SCALAR_FIELD = {0: 0, 1: 0, 2: 1, 3: 1}
G1_OF = {0: 0, 1: 2}
G2_OF = {0: 1, 1: 3}


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def hex_to_u64(h, width=None):
    a = np.frombuffer(bytes.fromhex(h), dtype="<u8").copy()
    if width:
        a = a.reshape(-1, width)
    return a


def u64_to_hex(a):
    return np.ascontiguousarray(a, dtype="<u8").tobytes().hex()


def int_to_limbs(v, n=5):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def ints_to_limbs(vs, n=5):
    out = np.zeros((len(vs), n), dtype=np.uint64)
    for i, v in enumerate(vs):
        out[i] = int_to_limbs(v, n)
    return out


def limbs_to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype="<u8").tobytes(), "little")


def random_field_elems(n, field, seed, plain=False):
    """n uniform elements of the field as (n, 5) uint64 limbs.  Rejection sampling on 298-bit
    draws (the shape of ark-ff's Fp::rand); the limbs are returned as they are drawn, i.e. read as
    Montgomery representation by default (any canonical limb vector < p is a valid element)."""
    p = FIELD_P[field]
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.zeros((n, 5), dtype=np.uint64)
    todo = np.arange(n)
    pl = int_to_limbs(p)
    while len(todo):
        d = rng.integers(0, 2 ** 64, size=(len(todo), 5), dtype=np.uint64)
        d[:, 4] &= np.uint64((1 << (298 - 256)) - 1)
        # lexicographic compare d < p from the top limb down
        lt = np.zeros(len(todo), dtype=bool)
        eq = np.ones(len(todo), dtype=bool)
        for i in range(4, -1, -1):
            lt |= eq & (d[:, i] < pl[i])
            eq &= d[:, i] == pl[i]
        out[todo[lt]] = d[lt]
        todo = todo[~lt]
    return out


def csr_from_golden(g):
    return (np.array(g["ptr"], dtype=np.uint32), np.array(g["col"], dtype=np.uint32),
            hex_to_u64(g["val"], 5) if g["val"] else np.zeros((0, 5), dtype=np.uint64))


GM17_G2_KEYS = ("b_query", "h_gamma_z")
GM17_SINGLE = ("g_gamma_z", "h_gamma_z", "g_ab_gamma_z", "g_gamma2_z2")


def gm17_pk_from_golden(case):
    """golden gm17.json key -> dict of numpy arrays keyed like c_oracle.GM17_PK_FIELDS"""
    pid = case["pairing"]
    g1, g2 = G1_OF[pid], G2_OF[pid]
    pk = {k: hex_to_u64(v, POINT_LIMBS[g2 if k in GM17_G2_KEYS else g1]) for k, v in case["pk"].items()}
    for k in GM17_SINGLE:
        pk[k] = pk[k][0]
    return pk
