"""Host-side mirror of ark-poly-commit's KZG10 (kzg10::KZG10::{commit, open}) and of the dense-polynomial
operations under ark-marlin's prover, for the reference's Marlin configuration
(/root/reference/tests/mnt4_marlin.rs:68-94: MarlinSNARK<Fr, Fq, MarlinKZG10<E, DensePolynomial<Fr>>, ...>).

Names and argument meaning follow ark-poly-commit: `Powers { powers_of_g, powers_of_gamma_g }` is the committer
key, `commit(powers, polynomial, hiding_bound, rng)` returns the commitment and the blinding polynomial
(`Randomness`), `open(powers, polynomial, point, randomness)` returns `Proof { w, random_v }`.  Everything numeric
runs in libpcdgpu.so; the AHP round logic and the Fiat-Shamir sponge stay on the host side of the boundary."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from . import lib as L


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class Powers:
    """kzg10::Powers: the committer key, resident on the GPU (optionally with precomputed window tables)."""

    def __init__(self, ctx: L.Context, pairing: int, powers_of_g: np.ndarray, powers_of_gamma_g: np.ndarray,
                 precompute: bool = False):
        self.ctx, self.pairing = ctx, pairing
        self.curve = L.G1_OF[pairing]
        self.field = L.SCALAR_FIELD_OF[pairing]
        self.powers_of_g = L.Bases(ctx, self.curve, powers_of_g, precompute)
        self.powers_of_gamma_g = L.Bases(ctx, self.curve, powers_of_gamma_g, precompute)

    def size(self) -> int:
        return self.powers_of_g.n

    def close(self):
        self.powers_of_g.close()
        self.powers_of_gamma_g.close()


@dataclass
class Randomness:
    """kzg10::Randomness: the blinding polynomial's coefficients (Montgomery limbs), empty when not hiding."""

    blinding_polynomial: np.ndarray

    def is_hiding(self) -> bool:
        return self.blinding_polynomial.shape[0] > 0


@dataclass
class OpeningProof:
    """kzg10::Proof { w, random_v }"""

    w: np.ndarray
    random_v: Optional[np.ndarray]


def hiding_blinding_coefficients(hiding_bound: int) -> int:
    """coefficients of the blinding polynomial of a hiding commitment: degree hiding_bound + 1 -> hiding_bound + 2 draws"""
    return hiding_bound + 2


class KZG10:
    def __init__(self, ctx: L.Context):
        self.ctx = ctx

    def commit(self, powers: Powers, polynomial: np.ndarray, hiding_bound: Optional[int] = None,
               rng: Optional[Callable[[int], np.ndarray]] = None):
        """KZG10::commit.  polynomial: (n, 5) Montgomery coefficients, lowest degree first.  With a hiding bound the
        blinding polynomial is `Randomness::rand(hiding_bound, ..)`: degree
        `calculate_hiding_polynomial_degree(hiding_bound) = hiding_bound + 1`, i.e. hiding_bound + 2 coefficients
        (`P::rand(d, rng)` draws d + 1), drawn from rng(field) in order, lowest degree first (ark-poly-commit
        kzg10/data_structures.rs, recalled: upstream is not vendored) -- the draw COUNT matters for byte parity because
        every later draw of the caller's rng shifts with it.  Returns (commitment affine limbs, Randomness)."""
        poly = np.ascontiguousarray(polynomial, dtype=np.uint64).reshape(-1, 5)
        if poly.shape[0] > powers.size():
            raise ValueError("polynomial has %d coefficients for %d powers" % (poly.shape[0], powers.size()))
        if hiding_bound is not None:
            if rng is None:
                raise ValueError("a hiding commitment needs an rng")
            blind = np.stack([rng(powers.field) for _ in range(hiding_blinding_coefficients(hiding_bound))]).astype(np.uint64)
        else:
            blind = np.zeros((0, 5), dtype=np.uint64)
        out = np.zeros(L.AFFINE_LIMBS[powers.curve], dtype=np.uint64)
        self.ctx._check(self.ctx.lib.pcdgpu_kzg_commit(self.ctx.h, powers.powers_of_g.h, _vp(poly), poly.shape[0],
                                                       powers.powers_of_gamma_g.h, _vp(blind), blind.shape[0], _vp(out)))
        return out, Randomness(blind)

    def open(self, powers: Powers, polynomial: np.ndarray, point: np.ndarray, randomness: Randomness):
        """KZG10::open -> (OpeningProof, p(point)); point: Montgomery limbs."""
        poly = np.ascontiguousarray(polynomial, dtype=np.uint64).reshape(-1, 5)
        point = np.ascontiguousarray(point, dtype=np.uint64)
        blind = np.ascontiguousarray(randomness.blinding_polynomial, dtype=np.uint64).reshape(-1, 5)
        w = np.zeros(L.AFFINE_LIMBS[powers.curve], dtype=np.uint64)
        value = np.zeros(5, dtype=np.uint64)
        rv = np.zeros(5, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.pcdgpu_kzg_open(self.ctx.h, powers.powers_of_g.h, _vp(poly), poly.shape[0],
                                                     powers.powers_of_gamma_g.h, _vp(blind), blind.shape[0], _vp(point),
                                                     _vp(w), _vp(value), _vp(rv)))
        return OpeningProof(w, rv if blind.shape[0] else None), value


def poly_divide_linear(ctx: L.Context, field: int, coeffs: np.ndarray, z: np.ndarray):
    """(p - p(z)) / (X - z) and p(z) (DensePolynomial division by a linear factor, as KZG10::open needs it)."""
    c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 5)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    q = np.zeros((max(c.shape[0] - 1, 0), 5), dtype=np.uint64)
    e = np.zeros(5, dtype=np.uint64)
    ctx._check(ctx.lib.pcdgpu_poly_divide_linear(ctx.h, field, _vp(c), c.shape[0], _vp(z), _vp(q) if q.size else None,
                                                 _vp(e)))
    return q, e


def poly_mul(ctx: L.Context, field: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """DensePolynomial `&a * &b` through the evaluation domain."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 5)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 5)
    out = np.zeros((a.shape[0] + b.shape[0] - 1, 5), dtype=np.uint64)
    ctx._check(ctx.lib.pcdgpu_poly_mul(ctx.h, field, _vp(a), a.shape[0], _vp(b), b.shape[0], _vp(out)))
    return out
