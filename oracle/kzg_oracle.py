"""CPU restatement of ark-poly-commit's KZG10 commit / open and the dense-polynomial operations under them
(ark-poly-commit src/kzg10/mod.rs: KZG10::commit, KZG10::open, compute_witness_polynomial; ark-poly
DensePolynomial division / multiplication) -- the commitment scheme the reference's Marlin configuration runs on
(/root/reference/tests/mnt4_marlin.rs:68-94).  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU leg
may import this.  PARITY UNPINNED: ark-poly-commit is an un-vendored, un-pinned git dependency
(/root/reference/Cargo.toml:42) and the reference holds no KZG vectors; what pins this file is (1) polynomial
identities checked against naive big-int evaluation and (2) an SRS with a KNOWN trapdoor beta, gamma:
commit(p) must equal [p(beta) + gamma r(beta)] G and the opening witness [(p(beta) - p(z)) / (beta - z) + ...] G.

Polynomials are Python int coefficient lists (plain integers mod p), lowest degree first.  Group arithmetic goes
through the C++ oracle (c_oracle.msm / fixed_base_mul), itself pinned against the Python big-int oracle."""
from typing import List, Optional, Sequence, Tuple

import numpy as np

import c_oracle as co
import pcd_oracle as o

PAIRINGS = {0: o.MNT4, 1: o.MNT6}


def poly_eval(p_: int, coeffs: Sequence[int], z: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % p_
    return acc


def poly_divide_linear(p_: int, coeffs: Sequence[int], z: int) -> Tuple[List[int], int]:
    """(p - p(z)) / (X - z) by synthetic division and p(z) (compute_witness_polynomial divides p - p(z) by X - z;
    the remainder of p / (X - z) is p(z), so the quotients agree)."""
    n = len(coeffs)
    q = [0] * max(n - 1, 0)
    acc = 0
    for j in range(n - 1, 0, -1):
        acc = (coeffs[j] + z * acc) % p_
        q[j - 1] = acc
    return q, (coeffs[0] + z * acc) % p_ if n else 0


def poly_mul_naive(p_: int, a: Sequence[int], b: Sequence[int]) -> List[int]:
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % p_
    return out


def _limbs(vals) -> np.ndarray:
    b = b"".join(int(v).to_bytes(40, "little") for v in vals)
    return np.frombuffer(b, dtype="<u8").copy().reshape(-1, 5)


def setup(pairing_id: int, max_degree: int, beta: int, gamma: int, generator_limbs: np.ndarray, threads: int = 0):
    """KZG10::setup with a known trapdoor: powers_of_g[i] = beta^i G, powers_of_gamma_g[i] = gamma beta^i G
    (affine limb arrays, max_degree + 1 / max_degree + 2 points as upstream)."""
    p_ = PAIRINGS[pairing_id].fr.p
    g1 = 0 if pairing_id == 0 else 2
    pw = [pow(beta, i, p_) for i in range(max_degree + 2)]
    pg = co.fixed_base_mul(g1, generator_limbs, _limbs(pw[:max_degree + 1]), threads)
    pgg = co.fixed_base_mul(g1, generator_limbs, _limbs([gamma * x % p_ for x in pw]), threads)
    return pg, pgg


def commit(pairing_id: int, powers_of_g: np.ndarray, powers_of_gamma_g: np.ndarray, coeffs: Sequence[int],
           blinding: Optional[Sequence[int]] = None, threads: int = 1) -> np.ndarray:
    """KZG10::commit: MSM(powers_of_g, coeffs) + MSM(powers_of_gamma_g, blinding); affine limbs."""
    g1 = 0 if pairing_id == 0 else 2
    assert len(coeffs) <= powers_of_g.shape[0]
    parts = [co.msm(g1, powers_of_g[:len(coeffs)], _limbs(coeffs), threads)] if len(coeffs) else []
    if blinding:
        assert len(blinding) <= powers_of_gamma_g.shape[0]
        parts.append(co.msm(g1, powers_of_gamma_g[:len(blinding)], _limbs(blinding), threads))
    if not parts:
        return np.zeros(10, dtype=np.uint64)
    return co.point_sum(g1, np.stack(parts))


def open_(pairing_id: int, powers_of_g: np.ndarray, powers_of_gamma_g: np.ndarray, coeffs: Sequence[int], z: int,
          blinding: Optional[Sequence[int]] = None, threads: int = 1):
    """KZG10::open -> (w affine limbs, p(z), random_v or None)."""
    p_ = PAIRINGS[pairing_id].fr.p
    q, value = poly_divide_linear(p_, coeffs, z)
    rq, rv = (poly_divide_linear(p_, blinding, z) if blinding else ([], None))
    w = commit(pairing_id, powers_of_g, powers_of_gamma_g, q, rq if blinding else None, threads)
    return w, value, rv


def expected_commit_log(p_: int, beta: int, gamma: int, coeffs, blinding=None) -> int:
    return (poly_eval(p_, coeffs, beta) + (gamma * poly_eval(p_, blinding, beta) if blinding else 0)) % p_


def expected_open_log(p_: int, beta: int, gamma: int, coeffs, z: int, blinding=None) -> int:
    inv = pow((beta - z) % p_, -1, p_)
    v = (poly_eval(p_, coeffs, beta) - poly_eval(p_, coeffs, z)) * inv
    if blinding:
        v += gamma * (poly_eval(p_, blinding, beta) - poly_eval(p_, blinding, z)) * inv
    return v % p_
