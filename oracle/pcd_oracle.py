"""CPU oracle for the Groth16 proving path of arkworks-rs/pcd on the MNT4-298 / MNT6-298 cycle.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it, and there only as the checker.  The product path (``pcd_b200`` -> ``libpcdgpu.so``)
never imports this module and has no CPU fallback.

PARITY UNPINNED.  The arithmetic on the hot path does not live under ``/root/reference``: it
lives in un-vendored, un-pinned git dependencies (ark-groth16, ark-poly, ark-ec, ark-ff,
ark-relations -- /root/reference/Cargo.toml:16-42, no rev/tag; Cargo.lock git-ignored,
/root/reference/.gitignore:2) and the reference's own tests hold no golden vectors for this
path (they only assert ``verify == true/false``: /root/reference/tests/mnt4_groth16.rs:87,103,
117,119).  There is no Rust toolchain here, so the reference cannot be run either.  This file
therefore *restates the published algorithms* of those crates (arkworks ~v0.2-0.3, SURVEY.md
Appendix B) in plain Python big-int arithmetic, anchored on the reference's call sites:

  * ``IC::MainSNARK::prove`` / ``IC::HelpSNARK::prove``  /root/reference/src/ec_cycle_pcd/mod.rs:171,179
  * tiny default-circuit proves                          /root/reference/src/ec_cycle_pcd/data_structures.rs:139-143,343-350
  * type bindings ``Groth16<MNT4_298>``/``Groth16<MNT6_298>``  /root/reference/tests/mnt4_groth16.rs:23-30

What pins it instead:
  1. every constant below is checked algebraically in ``self_check()`` (primality, 2-adicity,
     root orders, curve orders, twist orders);
  2. fast algorithms are checked against naive ones (naive DFT, double-and-add MSM);
  3. the Groth16 prover is checked end to end with a known trapdoor (``groth16_trapdoor_check``);
  4. Python <-> C (oracle/c) <-> CUDA agree byte-for-byte on seeded inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field
from typing import Callable, List, Optional, Sequence, Tuple

# ----------------------------------------------------------------------------------------------
# Constants (all re-verified in self_check()).
# ----------------------------------------------------------------------------------------------
#: MNT4-298 Fr = MNT6-298 Fq.  2-adicity 34.  arkworks: ark-mnt4-298 fields/fr.rs
R4 = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137
#: MNT4-298 Fq = MNT6-298 Fr.  2-adicity 17, 7-adicity 2.  arkworks: ark-mnt4-298 fields/fq.rs
Q4 = 475922286169261325753349249653048451545124879242694725395555128576210262817955800483758081

LIMBS64 = 5
MONT_BITS = 64 * LIMBS64  # R = 2^320 (ark-ff Fp320, 5 x u64)
MODULUS_BITS = 298


@dataclass(frozen=True)
class PrimeFieldParams:
    name: str
    p: int
    generator: int  # multiplicative generator (arkworks GENERATOR)
    two_adicity: int
    small_subgroup_base: Optional[int] = None
    small_subgroup_adicity: Optional[int] = None

    @property
    def R(self) -> int:
        return (1 << MONT_BITS) % self.p

    @property
    def R2(self) -> int:
        return (1 << (2 * MONT_BITS)) % self.p

    @property
    def two_adic_root(self) -> int:
        return pow(self.generator, (self.p - 1) >> self.two_adicity, self.p)

    @property
    def inv64(self) -> int:
        return (-pow(self.p, -1, 1 << 64)) % (1 << 64)

    @property
    def inv32(self) -> int:
        return (-pow(self.p, -1, 1 << 32)) % (1 << 32)


FR4 = PrimeFieldParams("r4", R4, 10, 34)
FQ4 = PrimeFieldParams("q4", Q4, 17, 17, 7, 2)

# Curves.  Short Weierstrass y^2 = x^3 + a x + b.
MNT4_G1_A = 2
MNT4_G1_B = 423894536526684178289416011533888240029318103673896002803341544124054745019340795360841685
MNT4_G1_GEN = (
    60760244141852568949126569781626075788424196370144486719385562369396875346601926534016838,
    363732850702582978263902770815145784459747722357071843971107674179038674942891694705904306,
)
MNT6_G1_A = 11
MNT6_G1_B = 106700080510851735677967319632585352256454251201367587890185989362936000262606668469523074
FQ2_NONRESIDUE = 17  # Fq2 = F_q4[u]/(u^2 - 17)
FQ3_NONRESIDUE = 5  # Fq3 = F_r4[u]/(u^3 - 5)


# ----------------------------------------------------------------------------------------------
# Field towers.  Elements: int for Fp, tuple of ints for Fp2 / Fp3.
# ----------------------------------------------------------------------------------------------
class Fp:
    """Prime field; elements are Python ints in [0, p)."""

    degree = 1

    def __init__(self, p: int):
        self.p = p
        self.zero = 0
        self.one = 1
        self.order = p

    def add(self, a, b):
        return (a + b) % self.p

    def sub(self, a, b):
        return (a - b) % self.p

    def neg(self, a):
        return (-a) % self.p

    def mul(self, a, b):
        return (a * b) % self.p

    def sqr(self, a):
        return (a * a) % self.p

    def inv(self, a):
        if a % self.p == 0:
            raise ZeroDivisionError("inverse of zero")
        return pow(a, -1, self.p)

    def from_int(self, a: int):
        return a % self.p

    def is_zero(self, a):
        return a % self.p == 0

    def eq(self, a, b):
        return (a - b) % self.p == 0

    def pow(self, a, e: int):
        return pow(a, e, self.p)

    def coeffs(self, a) -> Tuple[int, ...]:
        return (a,)

    def from_coeffs(self, c):
        return c[0] % self.p


class Fp2:
    """F_p[u]/(u^2 - nr); elements are (c0, c1).  ark-ff Fp2 (models/fp2.rs)."""

    degree = 2

    def __init__(self, p: int, nr: int):
        self.p = p
        self.nr = nr
        self.zero = (0, 0)
        self.one = (1, 0)
        self.order = p * p

    def add(self, a, b):
        return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)

    def neg(self, a):
        return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.nr * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def sqr(self, a):
        return self.mul(a, a)

    def inv(self, a):
        p = self.p
        n = (a[0] * a[0] - self.nr * a[1] * a[1]) % p
        if n == 0:
            raise ZeroDivisionError("inverse of zero")
        ni = pow(n, -1, p)
        return (a[0] * ni % p, (-a[1]) * ni % p)

    def from_int(self, a: int):
        return (a % self.p, 0)

    def is_zero(self, a):
        return a[0] % self.p == 0 and a[1] % self.p == 0

    def eq(self, a, b):
        return self.is_zero(self.sub(a, b))

    def pow(self, a, e: int):
        r = self.one
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r

    def coeffs(self, a):
        return tuple(a)

    def from_coeffs(self, c):
        return (c[0] % self.p, c[1] % self.p)


class Fp3:
    """F_p[u]/(u^3 - nr); elements are (c0, c1, c2).  ark-ff Fp3 (models/fp3.rs)."""

    degree = 3

    def __init__(self, p: int, nr: int):
        self.p = p
        self.nr = nr
        self.zero = (0, 0, 0)
        self.one = (1, 0, 0)
        self.order = p ** 3

    def add(self, a, b):
        p = self.p
        return ((a[0] + b[0]) % p, (a[1] + b[1]) % p, (a[2] + b[2]) % p)

    def sub(self, a, b):
        p = self.p
        return ((a[0] - b[0]) % p, (a[1] - b[1]) % p, (a[2] - b[2]) % p)

    def neg(self, a):
        p = self.p
        return ((-a[0]) % p, (-a[1]) % p, (-a[2]) % p)

    def mul(self, a, b):
        p, nr = self.p, self.nr
        a0, a1, a2 = a
        b0, b1, b2 = b
        return (
            (a0 * b0 + nr * (a1 * b2 + a2 * b1)) % p,
            (a0 * b1 + a1 * b0 + nr * a2 * b2) % p,
            (a0 * b2 + a1 * b1 + a2 * b0) % p,
        )

    def sqr(self, a):
        return self.mul(a, a)

    def inv(self, a):
        # Solve via the adjugate: standard Fp3 inverse.
        p, nr = self.p, self.nr
        a0, a1, a2 = a
        t0 = (a0 * a0 - nr * a1 * a2) % p
        t1 = (nr * a2 * a2 - a0 * a1) % p
        t2 = (a1 * a1 - a0 * a2) % p
        n = (a0 * t0 + nr * (a2 * t1 + a1 * t2)) % p
        if n == 0:
            raise ZeroDivisionError("inverse of zero")
        ni = pow(n, -1, p)
        return (t0 * ni % p, t1 * ni % p, t2 * ni % p)

    def from_int(self, a: int):
        return (a % self.p, 0, 0)

    def is_zero(self, a):
        return all(c % self.p == 0 for c in a)

    def eq(self, a, b):
        return self.is_zero(self.sub(a, b))

    def pow(self, a, e: int):
        r = self.one
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r

    def coeffs(self, a):
        return tuple(a)

    def from_coeffs(self, c):
        return (c[0] % self.p, c[1] % self.p, c[2] % self.p)


def field_sqrt(F, a):
    """Tonelli-Shanks in any finite field ``F`` with ``F.order`` elements; None if non-square."""
    if F.is_zero(a):
        return F.zero
    q1 = F.order - 1
    if not F.eq(F.pow(a, q1 // 2), F.one):
        return None
    s, t = 0, q1
    while t % 2 == 0:
        s += 1
        t //= 2
    # deterministic non-residue search
    k = 2
    while True:
        z = F.from_coeffs(tuple([k] + [1] * (F.degree - 1))) if F.degree > 1 else F.from_int(k)
        if not F.eq(F.pow(z, q1 // 2), F.one) and not F.is_zero(z):
            break
        k += 1
    c = F.pow(z, t)
    x = F.pow(a, (t + 1) // 2)
    b = F.pow(a, t)
    m = s
    while not F.eq(b, F.one):
        i, b2 = 0, b
        while not F.eq(b2, F.one):
            b2 = F.sqr(b2)
            i += 1
        e = F.pow(c, 1 << (m - i - 1))
        x = F.mul(x, e)
        c = F.sqr(e)
        b = F.mul(b, c)
        m = i
    return x


# ----------------------------------------------------------------------------------------------
# Curves.  Affine points are (x, y) or None for infinity.
# ----------------------------------------------------------------------------------------------
@dataclass
class Curve:
    name: str
    F: object  # coordinate field
    a: object
    b: object
    order: int  # prime subgroup order r
    cofactor: int
    gen: Optional[tuple] = None
    scalar_field: Optional[PrimeFieldParams] = None

    def is_on_curve(self, P) -> bool:
        if P is None:
            return True
        F = self.F
        x, y = P
        rhs = F.add(F.add(F.mul(F.sqr(x), x), F.mul(self.a, x)), self.b)
        return F.eq(F.sqr(y), rhs)

    def neg(self, P):
        if P is None:
            return None
        return (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if F.eq(x1, x2):
            if F.eq(y1, y2):
                return self.double(P)
            return None
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.sqr(lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def double(self, P):
        F = self.F
        if P is None:
            return None
        x1, y1 = P
        if F.is_zero(y1):
            return None
        three = F.from_int(3)
        lam = F.mul(F.add(F.mul(three, F.sqr(x1)), self.a), F.inv(F.add(y1, y1)))
        x3 = F.sub(F.sqr(lam), F.add(x1, x1))
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    # Jacobian arithmetic for speed (no inversions); used by mul() and the MSMs.
    def _jdbl(self, P):
        F = self.F
        X, Y, Z = P
        if F.is_zero(Z) or F.is_zero(Y):
            return (F.one, F.one, F.zero)
        XX = F.sqr(X)
        YY = F.sqr(Y)
        YYYY = F.sqr(YY)
        ZZ = F.sqr(Z)
        S = F.mul(F.from_int(4), F.mul(X, YY))
        M = F.add(F.mul(F.from_int(3), XX), F.mul(self.a, F.sqr(ZZ)))
        X3 = F.sub(F.sqr(M), F.add(S, S))
        Y3 = F.sub(F.mul(M, F.sub(S, X3)), F.mul(F.from_int(8), YYYY))
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def _jadd(self, P, Q):
        F = self.F
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        if F.is_zero(Z1):
            return Q
        if F.is_zero(Z2):
            return P
        Z1Z1 = F.sqr(Z1)
        Z2Z2 = F.sqr(Z2)
        U1 = F.mul(X1, Z2Z2)
        U2 = F.mul(X2, Z1Z1)
        S1 = F.mul(F.mul(Y1, Z2), Z2Z2)
        S2 = F.mul(F.mul(Y2, Z1), Z1Z1)
        if F.eq(U1, U2):
            if F.eq(S1, S2):
                return self._jdbl(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.sqr(H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.sqr(Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def to_jac(self, P):
        F = self.F
        if P is None:
            return (F.one, F.one, F.zero)
        return (P[0], P[1], F.one)

    def to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.sqr(zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def mul(self, P, k: int):
        """[k]P by double-and-add (the naive definition the fast paths are checked against)."""
        if P is None or k == 0:
            return None
        if k < 0:
            return self.mul(self.neg(P), -k)
        F = self.F
        acc = (F.one, F.one, F.zero)
        base = self.to_jac(P)
        for bit in bin(k)[2:]:
            acc = self._jdbl(acc)
            if bit == "1":
                acc = self._jadd(acc, base)
        return self.to_affine(acc)

    def sum(self, pts):
        F = self.F
        acc = (F.one, F.one, F.zero)
        for P in pts:
            acc = self._jadd(acc, self.to_jac(P))
        return self.to_affine(acc)


FQ4_F = Fp(Q4)
FR4_F = Fp(R4)
FQ2_F = Fp2(Q4, FQ2_NONRESIDUE)
FQ3_F = Fp3(R4, FQ3_NONRESIDUE)


def _trace(q: int, order: int) -> int:
    return q + 1 - order


def _twist_order_fq2(q: int, t: int) -> int:
    # #E(F_{q^2}) = q^2 + 1 - t2 with t2 = t^2 - 2q; the quadratic twist over F_{q^2} has +t2.
    t2 = t * t - 2 * q
    return q * q + 1 + t2


def _twist_order_fq3(q: int, t: int) -> int:
    t3 = t ** 3 - 3 * q * t
    return q ** 3 + 1 + t3


def _find_point(curve: Curve, start: int = 1):
    """Deterministic point: smallest x = (start + i, [1, ...]) with a square RHS; smaller-y root."""
    F = curve.F
    i = start
    while True:
        if F.degree == 1:
            x = F.from_int(i)
        else:
            x = F.from_coeffs(tuple([i] + [1] + [0] * (F.degree - 2)))
        rhs = F.add(F.add(F.mul(F.sqr(x), x), F.mul(curve.a, x)), curve.b)
        y = field_sqrt(F, rhs)
        if y is not None and not F.is_zero(y):
            ny = F.neg(y)
            if tuple(reversed(F.coeffs(ny))) < tuple(reversed(F.coeffs(y))):
                y = ny
            return (x, y)
        i += 1


def _build_curves():
    g1_4 = Curve("mnt4_g1", FQ4_F, MNT4_G1_A, MNT4_G1_B, R4, 1, MNT4_G1_GEN, FR4)
    t4 = _trace(Q4, R4)
    ord_g2_4 = _twist_order_fq2(Q4, t4)
    assert ord_g2_4 % R4 == 0
    g2_4 = Curve(
        "mnt4_g2",
        FQ2_F,
        (MNT4_G1_A * FQ2_NONRESIDUE % Q4, 0),
        (0, MNT4_G1_B * FQ2_NONRESIDUE % Q4),
        R4,
        ord_g2_4 // R4,
        None,
        FR4,
    )
    g1_6 = Curve("mnt6_g1", FR4_F, MNT6_G1_A, MNT6_G1_B, Q4, 1, None, FQ4)
    t6 = _trace(R4, Q4)
    ord_g2_6 = _twist_order_fq3(R4, t6)
    assert ord_g2_6 % Q4 == 0
    g2_6 = Curve(
        "mnt6_g2",
        FQ3_F,
        (0, 0, MNT6_G1_A),
        (MNT6_G1_B * FQ3_NONRESIDUE % R4, 0, 0),
        Q4,
        ord_g2_6 // Q4,
        None,
        FQ4,
    )
    return g1_4, g2_4, g1_6, g2_6


MNT4_G1, MNT4_G2, MNT6_G1, MNT6_G2 = _build_curves()

# Generators for the groups arkworks' generator could not be recalled/verified for.  These are
# DERIVED here (deterministic search + cofactor clearing), not arkworks' constants; the prover
# never uses a generator (keys are inputs), only the test key generator does.
_GEN_CACHE = {}


def generator(curve: Curve):
    if curve.gen is not None:
        return curve.gen
    if curve.name not in _GEN_CACHE:
        P = _find_point(curve)
        G = curve.mul(P, curve.cofactor)
        assert G is not None and curve.is_on_curve(G)
        assert curve.mul(G, curve.order) is None
        _GEN_CACHE[curve.name] = G
    return _GEN_CACHE[curve.name]


@dataclass(frozen=True)
class Pairing:
    """The SNARK's view of one curve of the cycle (ark-ec PairingEngine, without the pairing)."""

    name: str
    fr: PrimeFieldParams
    g1: Curve
    g2: Curve


MNT4 = Pairing("mnt4_298", FR4, MNT4_G1, MNT4_G2)
MNT6 = Pairing("mnt6_298", FQ4, MNT6_G1, MNT6_G2)


# ----------------------------------------------------------------------------------------------
# Deterministic PRNG shared with the C oracle and the CUDA test generators (SplitMix64).
# Field sampling mirrors ark-ff's ``Fp::rand`` structure: draw five u64 limbs,
# mask to 298 bits, reject if >= p.  (ark_std::test_rng() itself is ChaCha; the RNG stays on the
# Rust side of the ABI, so only the *shape* of the sampler matters here.)
# ----------------------------------------------------------------------------------------------
M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & M64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    def field(self, p: int) -> int:
        while True:
            v = 0
            for i in range(LIMBS64):
                v |= self.next_u64() << (64 * i)
            v &= (1 << MODULUS_BITS) - 1
            if v < p:
                return v


# ----------------------------------------------------------------------------------------------
# Memory encodings (ark-ff BigInteger320 / Fp320 in-memory layout; ark-serialize canonical form).
# ----------------------------------------------------------------------------------------------
def fp_to_mont_bytes(a: int, fp: PrimeFieldParams) -> bytes:
    """40 bytes: little-endian limbs of a*R mod p (ark-ff Fp320's in-memory representation)."""
    return ((a * fp.R) % fp.p).to_bytes(40, "little")


def fp_from_mont_bytes(b: bytes, fp: PrimeFieldParams) -> int:
    v = int.from_bytes(b, "little")
    assert v < fp.p, "non-canonical Montgomery limbs"
    return (v * pow(fp.R, -1, fp.p)) % fp.p


def int_to_repr_bytes(a: int) -> bytes:
    """40 bytes: little-endian limbs of the plain integer (ark-ff ``into_repr()``)."""
    return a.to_bytes(40, "little")


def serialize_fp(a: int, flags: int = 0) -> bytes:
    """ark-serialize CanonicalSerialize for a 298-bit field element: 38 bytes LE, flags on top."""
    out = bytearray(a.to_bytes(38, "little"))
    out[-1] |= flags
    return bytes(out)


def _lex_larger(F, y) -> bool:
    """ark-ec compressed flag: is y the larger of {y, -y}; extension fields compare the highest
    coefficient first."""
    ny = F.neg(y)
    return tuple(reversed(F.coeffs(y))) > tuple(reversed(F.coeffs(ny)))


def serialize_point(curve: Curve, P) -> bytes:
    """Compressed short-Weierstrass affine point: x, bit7 = y is the larger root, bit6 = infinity."""
    F = curve.F
    if P is None:
        cs = [0] * F.degree
        out = b"".join(serialize_fp(0) for _ in cs[:-1]) + serialize_fp(0, 0x40)
        return out
    x, y = P
    cs = F.coeffs(x)
    flag = 0x80 if _lex_larger(F, y) else 0
    return b"".join(serialize_fp(c) for c in cs[:-1]) + serialize_fp(cs[-1], flag)


# ----------------------------------------------------------------------------------------------
# Evaluation domains and FFT (ark-poly GeneralEvaluationDomain / Radix2EvaluationDomain; B.3).
# ----------------------------------------------------------------------------------------------
@dataclass
class Domain:
    fp: PrimeFieldParams
    size: int
    log2: Optional[int]  # None for mixed radix
    omega: int  # group generator
    kind: str  # "radix2" | "mixed"

    @property
    def p(self):
        return self.fp.p

    @property
    def coset_gen(self):
        return self.fp.generator


def domain_new(fp: PrimeFieldParams, min_size: int) -> Domain:
    """``GeneralEvaluationDomain::new(m)``: radix-2 if log2(next_pow2(m)) <= TWO_ADICITY, else the
    smallest q^a * 2^b >= m when the field has a small subgroup (q=7, adicity 2 on q4)."""
    size = 1
    log = 0
    while size < max(min_size, 1):
        size <<= 1
        log += 1
    if log <= fp.two_adicity:
        omega = pow(fp.two_adic_root, 1 << (fp.two_adicity - log), fp.p)
        return Domain(fp, size, log, omega, "radix2")
    if fp.small_subgroup_base is None:
        raise ValueError("domain too large for field %s" % fp.name)
    best = None
    q = fp.small_subgroup_base
    for a in range(fp.small_subgroup_adicity + 1):
        for b in range(fp.two_adicity + 1):
            s = (q ** a) << b
            if s >= min_size and (best is None or s < best[0]):
                best = (s, a, b)
    if best is None:
        raise ValueError("domain too large for field %s" % fp.name)
    s, a, b = best
    full = (q ** fp.small_subgroup_adicity) << fp.two_adicity
    large_root = pow(fp.generator, (fp.p - 1) // full, fp.p)
    omega = pow(large_root, full // s, fp.p)
    return Domain(fp, s, None, omega, "mixed")


def domain_mixed(fp: PrimeFieldParams, a: int, b: int) -> Domain:
    """The mixed-radix domain of size q^a 2^b explicitly (tests: small sizes that domain_new would
    serve with radix 2)."""
    q = fp.small_subgroup_base
    assert q is not None and 0 <= a <= fp.small_subgroup_adicity and 0 <= b <= fp.two_adicity
    s = (q ** a) << b
    full = (q ** fp.small_subgroup_adicity) << fp.two_adicity
    large_root = pow(fp.generator, (fp.p - 1) // full, fp.p)
    return Domain(fp, s, None if a else b, pow(large_root, full // s, fp.p), "mixed" if a else "radix2")


def dft_naive(vals: Sequence[int], omega: int, p: int) -> List[int]:
    """out[i] = sum_j in[j] * omega^(i*j): the definition every NTT is checked against."""
    n = len(vals)
    pw = [1] * n
    for i in range(1, n):
        pw[i] = pw[i - 1] * omega % p
    return [sum(vals[j] * pw[(i * j) % n] for j in range(n)) % p for i in range(n)]


def _bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def fft_radix2(vals: Sequence[int], omega: int, p: int) -> List[int]:
    """In-order radix-2 Cooley-Tukey; natural in, natural out (same map as dft_naive)."""
    n = len(vals)
    if n == 1:
        return list(vals)
    bits = n.bit_length() - 1
    assert 1 << bits == n
    a = [vals[_bitrev(i, bits)] for i in range(n)]
    m = 1
    while m < n:
        wm = pow(omega, n // (2 * m), p)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % p
                u = a[k + j]
                a[k + j] = (u + t) % p
                a[k + j + m] = (u - t) % p
                w = w * wm % p
        m <<= 1
    return a


def domain_fft(d: Domain, vals: Sequence[int]) -> List[int]:
    v = list(vals) + [0] * (d.size - len(vals))
    assert len(v) == d.size
    if d.kind == "radix2":
        return fft_radix2(v, d.omega, d.p)
    return dft_naive(v, d.omega, d.p)


def domain_ifft(d: Domain, vals: Sequence[int]) -> List[int]:
    v = list(vals) + [0] * (d.size - len(vals))
    p = d.p
    oi = pow(d.omega, -1, p)
    out = fft_radix2(v, oi, p) if d.kind == "radix2" else dft_naive(v, oi, p)
    ninv = pow(d.size, -1, p)
    return [x * ninv % p for x in out]


def domain_coset_fft(d: Domain, vals: Sequence[int]) -> List[int]:
    p, g = d.p, d.coset_gen
    v = list(vals) + [0] * (d.size - len(vals))
    pw = 1
    for i in range(len(v)):
        v[i] = v[i] * pw % p
        pw = pw * g % p
    return domain_fft(d, v)


def domain_coset_ifft(d: Domain, vals: Sequence[int]) -> List[int]:
    p = d.p
    v = domain_ifft(d, vals)
    gi = pow(d.coset_gen, -1, p)
    pw = 1
    for i in range(len(v)):
        v[i] = v[i] * pw % p
        pw = pw * gi % p
    return v


# ----------------------------------------------------------------------------------------------
# Variable-base MSM (ark-ec VariableBaseMSM::multi_scalar_mul).
# ----------------------------------------------------------------------------------------------
def msm_naive(curve: Curve, bases: Sequence, scalars: Sequence[int]):
    F = curve.F
    acc = (F.one, F.one, F.zero)
    for P, k in zip(bases, scalars):
        if P is None or k == 0:
            continue
        acc = curve._jadd(acc, curve.to_jac(curve.mul(P, k)))
    return curve.to_affine(acc)


def ark_window_size(n: int) -> int:
    """c = 3 if N < 32 else floor(0.69 * ceil(log2 N)) + 2  (ark-ec msm/variable_base.rs)."""
    if n < 32:
        return 3
    lg = (n - 1).bit_length()  # ceil(log2 n) for n >= 2
    return (lg * 69) // 100 + 2


def msm_pippenger(curve: Curve, bases: Sequence, scalars: Sequence[int], c: Optional[int] = None):
    """arkworks-shaped Pippenger: truncate to the shorter input, drop zero scalars, scalars equal
    to one are added directly in window 0 only, unsigned c-bit digits into 2^c - 1 buckets,
    running-sum reduction, windows combined high -> low with c doublings each."""
    F = curve.F
    n = min(len(bases), len(scalars))
    pairs = [(bases[i], scalars[i]) for i in range(n) if scalars[i] != 0]
    if c is None:
        c = ark_window_size(n)
    num_bits = MODULUS_BITS
    zero = (F.one, F.one, F.zero)
    window_sums = []
    for w_start in range(0, num_bits, c):
        res = zero
        buckets = [zero] * ((1 << c) - 1)
        for P, k in pairs:
            if k == 1:
                if w_start == 0:
                    res = curve._jadd(res, curve.to_jac(P))
                continue
            d = (k >> w_start) & ((1 << c) - 1)
            if d != 0:
                buckets[d - 1] = curve._jadd(buckets[d - 1], curve.to_jac(P))
        running = zero
        for b in reversed(buckets):
            running = curve._jadd(running, b)
            res = curve._jadd(res, running)
        window_sums.append(res)
    total = window_sums[-1]
    for ws in reversed(window_sums[:-1]):
        for _ in range(c):
            total = curve._jdbl(total)
        total = curve._jadd(total, ws)
    return curve.to_affine(total)


# ----------------------------------------------------------------------------------------------
# R1CS (ark-relations ConstraintSystem::to_matrices) and the QAP witness map (ark-groth16
# r1cs_to_qap.rs).
# ----------------------------------------------------------------------------------------------
@dataclass
class R1CS:
    """Matrices as row lists of (coeff, col); columns: instance (col 0 is the constant 1), then
    witness -- exactly the layout of ark-relations' ``ConstraintMatrices``."""

    fp: PrimeFieldParams
    num_inputs: int  # includes the leading 1
    num_witness: int
    A: List[List[Tuple[int, int]]]
    B: List[List[Tuple[int, int]]]
    C: List[List[Tuple[int, int]]]

    @property
    def num_constraints(self) -> int:
        return len(self.A)

    @property
    def num_vars(self) -> int:
        return self.num_inputs + self.num_witness

    def is_satisfied(self, z: Sequence[int]) -> bool:
        p = self.fp.p
        for ra, rb, rc in zip(self.A, self.B, self.C):
            a = sum(c * z[j] for c, j in ra) % p
            b = sum(c * z[j] for c, j in rb) % p
            cc = sum(c * z[j] for c, j in rc) % p
            if (a * b - cc) % p:
                return False
        return True


def synthetic_r1cs(fp: PrimeFieldParams, num_constraints: int, num_inputs: int = 2, seed: int = 20261017,
                   bitlike: float = 0.0):
    """Satisfiable synthetic R1CS of the shape in SURVEY.md 8(d): row i of A and B has 1-4
    non-zeros over earlier variables with coefficients from {1, -1, 2^j, uniform}; C_i selects a
    fresh witness variable set to <A_i,z>*<B_i,z>.  ``bitlike`` = fraction of rows that instead
    constrain a fresh boolean (b*(b-1)=0 written as b*b=b), giving witness-like 0/1 assignments.
    Returns (r1cs, z) with z = instance || witness, z[0] = 1."""
    rng = SplitMix64(seed)
    p = fp.p
    z = [1] + [rng.field(p) for _ in range(num_inputs - 1)]
    A, B, C = [], [], []
    for i in range(num_constraints):
        nv = len(z)
        if bitlike and (rng.next_u64() % 1000) < int(bitlike * 1000):
            bit = rng.next_u64() & 1
            z.append(bit)
            A.append([(1, nv)])
            B.append([(1, nv)])
            C.append([(1, nv)])
            continue

        def row():
            k = 1 + rng.next_u64() % 4
            out = []
            for _ in range(k):
                col = rng.next_u64() % nv
                sel = rng.next_u64() % 4
                if sel == 0:
                    coeff = 1
                elif sel == 1:
                    coeff = p - 1
                elif sel == 2:
                    coeff = pow(2, rng.next_u64() % 64, p)
                else:
                    coeff = rng.field(p)
                out.append((coeff, col))
            return out

        ra, rb = row(), row()
        a = sum(c * z[j] for c, j in ra) % p
        b = sum(c * z[j] for c, j in rb) % p
        z.append(a * b % p)
        A.append(ra)
        B.append(rb)
        C.append([(1, nv)])
    r1cs = R1CS(fp, num_inputs, len(z) - num_inputs, A, B, C)
    return r1cs, z


def witness_map(r1cs: R1CS, z: Sequence[int]) -> Tuple[List[int], Domain]:
    """``R1CStoQAP::witness_map``: returns the n coefficients of h and the domain."""
    fp = r1cs.fp
    p = fp.p
    m = r1cs.num_constraints
    d = domain_new(fp, m + r1cs.num_inputs)
    n = d.size
    a = [0] * n
    b = [0] * n
    c = [0] * n
    for i in range(m):
        a[i] = sum(co * z[j] for co, j in r1cs.A[i]) % p
        b[i] = sum(co * z[j] for co, j in r1cs.B[i]) % p
        c[i] = sum(co * z[j] for co, j in r1cs.C[i]) % p
    for j in range(r1cs.num_inputs):
        a[m + j] = z[j]
    a = domain_coset_fft(d, domain_ifft(d, a))
    b = domain_coset_fft(d, domain_ifft(d, b))
    c = domain_coset_fft(d, domain_ifft(d, c))
    zinv = pow((pow(d.coset_gen, n, p) - 1) % p, -1, p)
    ab = [((a[i] * b[i] - c[i]) * zinv) % p for i in range(n)]
    h = domain_coset_ifft(d, ab)
    return h, d


# ----------------------------------------------------------------------------------------------
# Groth16 (ark-groth16 generator.rs / prover.rs) with a KNOWN trapdoor so that
# every proof element has a known discrete log.
# ----------------------------------------------------------------------------------------------
@dataclass
class Groth16PK:
    pairing: Pairing
    alpha_g1: tuple
    beta_g1: tuple
    delta_g1: tuple
    beta_g2: tuple
    delta_g2: tuple
    a_query: list  # G1, one per variable
    b_g1_query: list  # G1
    b_g2_query: list  # G2
    h_query: list  # G1, n - 1
    l_query: list  # G1, one per witness variable
    # test-only: the trapdoor and per-variable QAP evaluations
    trapdoor: dict = dc_field(default_factory=dict)


def _lagrange_at(d: Domain, tau: int) -> List[int]:
    """L_i(tau) for the domain (ark-poly evaluate_all_lagrange_coefficients):
    L_i(tau) = Z(tau) * w^i / (n * (tau - w^i)), with one batched inversion."""
    p, n = d.p, d.size
    zt = (pow(tau, n, p) - 1) % p
    assert zt != 0
    ninv = pow(n, -1, p)
    ws, dens = [], []
    w = 1
    for i in range(n):
        ws.append(w)
        dens.append((tau - w) % p)
        w = w * d.omega % p
    pref = [1] * (n + 1)
    for i in range(n):
        pref[i + 1] = pref[i] * dens[i] % p
    inv_all = pow(pref[n], -1, p)
    out = [0] * n
    c = zt * ninv % p
    for i in range(n - 1, -1, -1):
        out[i] = c * ws[i] % p * (inv_all * pref[i] % p) % p
        inv_all = inv_all * dens[i] % p
    return out


def groth16_setup_scalars(pairing: Pairing, r1cs: R1CS, seed: int = 7) -> dict:
    """The discrete logs of every proving-key element for a known trapdoor (test infrastructure):
    At/Bt/Ct = the QAP polynomials of each variable evaluated at tau, h_sc[i] = tau^i Z(tau)/delta,
    l_sc[j] = (beta At + alpha Bt + Ct)/delta for the witness variables."""
    fp = pairing.fr
    assert fp is r1cs.fp
    p = fp.p
    rng = SplitMix64(seed)
    alpha, beta, gamma, delta, tau = (rng.field(p) or 1 for _ in range(5))
    m = r1cs.num_constraints
    d = domain_new(fp, m + r1cs.num_inputs)
    n = d.size
    L = _lagrange_at(d, tau)
    nv = r1cs.num_vars
    At = [0] * nv
    Bt = [0] * nv
    Ct = [0] * nv
    for i in range(m):
        for co, j in r1cs.A[i]:
            At[j] = (At[j] + co * L[i]) % p
        for co, j in r1cs.B[i]:
            Bt[j] = (Bt[j] + co * L[i]) % p
        for co, j in r1cs.C[i]:
            Ct[j] = (Ct[j] + co * L[i]) % p
    for j in range(r1cs.num_inputs):
        At[j] = (At[j] + L[m + j]) % p
    zt = (pow(tau, n, p) - 1) % p
    dinv = pow(delta, -1, p)
    h_sc = []
    cur = zt * dinv % p
    for i in range(n - 1):
        h_sc.append(cur)
        cur = cur * tau % p
    l_sc = [(beta * At[j] + alpha * Bt[j] + Ct[j]) % p * dinv % p for j in range(r1cs.num_inputs, nv)]
    return dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, tau=tau, At=At, Bt=Bt, Ct=Ct,
                h_sc=h_sc, l_sc=l_sc)


def groth16_setup(pairing: Pairing, r1cs: R1CS, seed: int = 7) -> Groth16PK:
    """``generate_random_parameters`` restated with the trapdoor kept (test infrastructure)."""
    t = groth16_setup_scalars(pairing, r1cs, seed)
    G1, G2 = pairing.g1, pairing.g2
    g1, g2 = generator(G1), generator(G2)
    return Groth16PK(
        pairing,
        G1.mul(g1, t["alpha"]),
        G1.mul(g1, t["beta"]),
        G1.mul(g1, t["delta"]),
        G2.mul(g2, t["beta"]),
        G2.mul(g2, t["delta"]),
        [G1.mul(g1, x) for x in t["At"]],
        [G1.mul(g1, x) for x in t["Bt"]],
        [G2.mul(g2, x) for x in t["Bt"]],
        [G1.mul(g1, x) for x in t["h_sc"]],
        [G1.mul(g1, x) for x in t["l_sc"]],
        t,
    )


def groth16_prove(pk: Groth16PK, r1cs: R1CS, z: Sequence[int], r: int, s: int, msm=msm_pippenger):
    """``create_proof_with_reduction``, r and s supplied by the caller.  Returns
    (A in G1, B in G2, C in G1) as affine points."""
    pairing = pk.pairing
    G1, G2 = pairing.g1, pairing.g2
    p = pairing.fr.p
    h, _ = witness_map(r1cs, z)
    h_acc = msm(G1, pk.h_query, h)  # lengths n-1 vs n: truncated to the shorter
    aux = list(z[r1cs.num_inputs:])
    l_acc = msm(G1, pk.l_query, aux)
    assignment = list(z[1:])
    r_delta = G1.mul(pk.delta_g1, r)
    g_a = G1.sum([r_delta, pk.a_query[0], msm(G1, pk.a_query[1:], assignment), pk.alpha_g1])
    s_delta = G1.mul(pk.delta_g1, s)
    g1_b = G1.sum([s_delta, pk.b_g1_query[0], msm(G1, pk.b_g1_query[1:], assignment), pk.beta_g1])
    g2_b = G2.sum([G2.mul(pk.delta_g2, s), pk.b_g2_query[0], msm(G2, pk.b_g2_query[1:], assignment),
                   pk.beta_g2])
    rs_delta = G1.mul(pk.delta_g1, r * s % p)
    g_c = G1.sum([G1.mul(g_a, s), G1.mul(g1_b, r), G1.neg(rs_delta), l_acc, h_acc])
    return g_a, g2_b, g_c


def groth16_trapdoor_check(pk: Groth16PK, r1cs: R1CS, z: Sequence[int], r: int, s: int, proof) -> bool:
    """Check a proof through its known discrete logs (no pairing needed): A = [alpha + sum z_i
    A_i(tau) + r delta]G1, B likewise in G2, C = [(sum_aux z_j l_j + h(tau) Z(tau))/delta + s a +
    r b - r s delta]G1, with h(tau) taken from the QAP identity (A.z)(B.z) - (C.z) = h Z."""
    t = pk.trapdoor
    pairing = pk.pairing
    p = pairing.fr.p
    G1, G2 = pairing.g1, pairing.g2
    At, Bt, Ct = t["At"], t["Bt"], t["Ct"]
    az = sum(zi * x for zi, x in zip(z, At)) % p
    bz = sum(zi * x for zi, x in zip(z, Bt)) % p
    cz = sum(zi * x for zi, x in zip(z, Ct)) % p
    a_log = (t["alpha"] + az + r * t["delta"]) % p
    b_log = (t["beta"] + bz + s * t["delta"]) % p
    dinv = pow(t["delta"], -1, p)
    ni = r1cs.num_inputs
    l_part = sum(z[ni + j] * t["l_sc"][j] for j in range(r1cs.num_witness)) % p
    h_part = (az * bz - cz) % p * dinv % p  # = h(tau) Z(tau) / delta when the R1CS is satisfied
    c_log = (l_part + h_part + s * a_log + r * b_log - r * s % p * t["delta"]) % p
    A, B, C = proof
    return (
        A == G1.mul(generator(G1), a_log)
        and B == G2.mul(generator(G2), b_log)
        and C == G1.mul(generator(G1), c_log)
    )


def serialize_proof(pairing: Pairing, proof) -> bytes:
    """ark-groth16 ``Proof`` canonical bytes: a || b || c compressed (152 B MNT4, 190 B MNT6)."""
    A, B, C = proof
    return serialize_point(pairing.g1, A) + serialize_point(pairing.g2, B) + serialize_point(pairing.g1, C)


# ----------------------------------------------------------------------------------------------
# GM17 (ark-gm17 r1cs_to_sap.rs / generator.rs / prover.rs), the second SNARK the
# reference plugs into ECCyclePCD (/root/reference/tests/mnt4_gm17.rs:27-28, mnt4_mix_*.rs).  Restated
# from the published construction (Groth-Maller 2017 over square arithmetic programs, in the shape of
# libsnark's r1cs_se_ppzksnark that ark-gm17 follows); PARITY UNPINNED like everything else here.  The
# key is generated with a KNOWN trapdoor so a proof can be checked through its discrete logarithms.
# ----------------------------------------------------------------------------------------------
def _row_eval(row, z, p):
    return sum(co * z[j] for co, j in row) % p


def sap_extend_assignment(r1cs: R1CS, z: Sequence[int]) -> List[int]:
    """``R1CStoSAP::witness_map``, first half: the SAP assignment = z || (<A_i,z> - <B_i,z>)^2 for every
    constraint || (z_i - 1)^2 for every public input i >= 1."""
    p = r1cs.fp.p
    full = list(z)
    for ra, rb in zip(r1cs.A, r1cs.B):
        full.append(pow((_row_eval(ra, z, p) - _row_eval(rb, z, p)) % p, 2, p))
    for i in range(1, r1cs.num_inputs):
        full.append(pow((z[i] - 1) % p, 2, p))
    return full


def sap_domain(r1cs: R1CS) -> Domain:
    return domain_new(r1cs.fp, 2 * r1cs.num_constraints + 2 * (r1cs.num_inputs - 1) + 1)


def sap_witness_map(r1cs: R1CS, z: Sequence[int], d1: int, d2: int):
    """``R1CStoSAP::witness_map``: returns (full SAP assignment, the n + 1 coefficients of
    H = ((u + d1 Z)^2 - (c + d2 Z)) / Z, domain).  SAP rows: constraint i gives rows 2i: (A_i + B_i)^2 =
    4 C_i + e_i and 2i + 1: (A_i - B_i)^2 = e_i; then 1^2 = 1; then for each input (x + 1)^2 = 4 x + e', (x - 1)^2 = e'."""
    fp = r1cs.fp
    p = fp.p
    m, ni = r1cs.num_constraints, r1cs.num_inputs
    full = sap_extend_assignment(r1cs, z)
    d = sap_domain(r1cs)
    n = d.size
    off = 2 * m
    ev1 = ni + r1cs.num_witness          # first per-constraint extra variable
    ev2 = ev1 + m - 1                    # extra variable of input i is at ev2 + i
    a = [0] * n
    c = [0] * n
    for i in range(m):
        az, bz, cz = (_row_eval(r, z, p) for r in (r1cs.A[i], r1cs.B[i], r1cs.C[i]))
        a[2 * i] = (az + bz) % p
        a[2 * i + 1] = (az - bz) % p
        c[2 * i] = (4 * cz + full[ev1 + i]) % p
        c[2 * i + 1] = full[ev1 + i]
    a[off] = 1
    c[off] = 1
    for i in range(1, ni):
        a[off + 2 * i - 1] = (z[i] + 1) % p
        a[off + 2 * i] = (z[i] - 1) % p
        c[off + 2 * i - 1] = (4 * z[i] + full[ev2 + i]) % p
        c[off + 2 * i] = full[ev2 + i]
    a = domain_ifft(d, a)
    h = [(2 * d1 * x) % p for x in a]
    h[0] = (h[0] - d2 - d1 * d1) % p
    h.append(d1 * d1 % p)
    a = domain_coset_fft(d, a)
    c = domain_coset_fft(d, domain_ifft(d, c))
    zinv = pow((pow(d.coset_gen, n, p) - 1) % p, -1, p)
    aa = domain_coset_ifft(d, [((a[i] * a[i] - c[i]) * zinv) % p for i in range(n)])
    for i in range(n - 1):
        h[i] = (h[i] + aa[i]) % p
    assert aa[n - 1] == 0, "SAP not satisfied: quotient degree too high"
    return full, h, d


@dataclass
class GM17PK:
    pairing: Pairing
    # vk
    h_g2: tuple
    g_alpha_g1: tuple
    h_beta_g2: tuple
    g_gamma_g1: tuple
    h_gamma_g2: tuple
    vk_query: list          # G1, one per public input (incl. the constant)
    # pk
    a_query: list           # G1, one per SAP variable: gamma At_i
    b_query: list           # G2, one per SAP variable: gamma At_i
    c_query_1: list         # G1, non-input SAP variables: gamma^2 Ct_i + (alpha + beta) gamma At_i
    c_query_2: list         # G1, one per SAP variable: 2 gamma^2 Z At_i
    g_gamma_z: tuple
    h_gamma_z: tuple
    g_ab_gamma_z: tuple
    g_gamma2_z2: tuple
    g_gamma2_z_t: list      # G1, n + 1 powers: gamma^2 Z t^i
    trapdoor: dict = dc_field(default_factory=dict)


def gm17_setup_scalars(pairing: Pairing, r1cs: R1CS, seed: int = 11) -> dict:
    """``R1CStoSAP::instance_map_with_evaluation`` + the exponents of ``generate_parameters`` for a known
    trapdoor (t, alpha, beta, gamma); generators are the fixed group generators (upstream draws random ones)."""
    fp = pairing.fr
    assert fp is r1cs.fp
    p = fp.p
    rng = SplitMix64(seed)
    alpha, beta, gamma, t = (rng.field(p) or 1 for _ in range(4))
    m, ni = r1cs.num_constraints, r1cs.num_inputs
    d = sap_domain(r1cs)
    n = d.size
    u = _lagrange_at(d, t)
    nsap = ni + r1cs.num_witness + m + (ni - 1)
    ev1 = ni + r1cs.num_witness
    ev2 = ev1 + m - 1
    off = 2 * m
    At = [0] * nsap
    Ct = [0] * nsap
    for i in range(m):
        uadd, usub = (u[2 * i] + u[2 * i + 1]) % p, (u[2 * i] - u[2 * i + 1]) % p
        for co, j in r1cs.A[i]:
            At[j] = (At[j] + uadd * co) % p
        for co, j in r1cs.B[i]:
            At[j] = (At[j] + usub * co) % p
        for co, j in r1cs.C[i]:
            Ct[j] = (Ct[j] + 4 * u[2 * i] * co) % p
        Ct[ev1 + i] = (Ct[ev1 + i] + uadd) % p
    At[0] = (At[0] + u[off]) % p
    Ct[0] = (Ct[0] + u[off]) % p
    for i in range(1, ni):
        u1, u2 = u[off + 2 * i - 1], u[off + 2 * i]
        At[i] = (At[i] + u1 + u2) % p
        At[0] = (At[0] + u1 - u2) % p
        Ct[i] = (Ct[i] + 4 * u1) % p
        Ct[ev2 + i] = (Ct[ev2 + i] + u1 + u2) % p
    zt = (pow(t, n, p) - 1) % p
    ab = (alpha + beta) % p
    g2 = gamma * gamma % p
    return dict(
        alpha=alpha, beta=beta, gamma=gamma, t=t, zt=zt, At=At, Ct=Ct, n=n,
        a_sc=[gamma * x % p for x in At],
        c1_sc=[(g2 * Ct[j] + ab * gamma % p * At[j]) % p for j in range(ni, nsap)],
        c2_sc=[2 * g2 % p * zt % p * x % p for x in At],
        vk_sc=[(gamma * Ct[j] + ab * At[j]) % p for j in range(ni)],
        gzt_sc=[g2 * zt % p * pow(t, i, p) % p for i in range(n + 1)],
    )


def gm17_setup(pairing: Pairing, r1cs: R1CS, seed: int = 11) -> GM17PK:
    t = gm17_setup_scalars(pairing, r1cs, seed)
    p = pairing.fr.p
    G1, G2 = pairing.g1, pairing.g2
    g, h = generator(G1), generator(G2)
    gam, zt, ab = t["gamma"], t["zt"], (t["alpha"] + t["beta"]) % p
    return GM17PK(
        pairing, h, G1.mul(g, t["alpha"]), G2.mul(h, t["beta"]), G1.mul(g, gam), G2.mul(h, gam),
        [G1.mul(g, x) for x in t["vk_sc"]],
        [G1.mul(g, x) for x in t["a_sc"]],
        [G2.mul(h, x) for x in t["a_sc"]],
        [G1.mul(g, x) for x in t["c1_sc"]],
        [G1.mul(g, x) for x in t["c2_sc"]],
        G1.mul(g, gam * zt % p), G2.mul(h, gam * zt % p), G1.mul(g, ab * gam % p * zt % p),
        G1.mul(g, gam * gam % p * zt % p * zt % p),
        [G1.mul(g, x) for x in t["gzt_sc"]],
        t,
    )


def gm17_prove(pk: GM17PK, r1cs: R1CS, z: Sequence[int], d1: int, d2: int, r: int, msm=msm_pippenger):
    """``create_proof`` of ark-gm17 (d1, d2, r supplied by the caller, drawn in this order upstream).  The
    input / aux splits of upstream's MSMs are not reproduced: a sum of MSMs over a partition of the
    index set is the MSM over the whole set, and proofs are compared after into_affine()."""
    pairing = pk.pairing
    G1, G2 = pairing.g1, pairing.g2
    p = pairing.fr.p
    ni = r1cs.num_inputs
    full, h, _ = sap_witness_map(r1cs, z, d1, d2)
    rest = full[1:]
    aux = full[ni:]
    g_a = G1.sum([G1.mul(pk.g_gamma_z, r), pk.a_query[0], G1.mul(pk.g_gamma_z, d1), msm(G1, pk.a_query[1:], rest)])
    g_b = G2.sum([G2.mul(pk.h_gamma_z, r), pk.b_query[0], G2.mul(pk.h_gamma_z, d1), msm(G2, pk.b_query[1:], rest)])
    c1_acc = msm(G1, pk.c_query_1, aux)
    c2_acc = msm(G1, pk.c_query_2[1:], rest)
    g_acc = msm(G1, pk.g_gamma2_z_t, h)
    g_c = G1.sum([
        c1_acc,
        G1.mul(pk.g_gamma2_z2, r * r % p),
        G1.mul(pk.g_ab_gamma_z, r),
        G1.mul(pk.g_ab_gamma_z, d1),
        G1.mul(pk.c_query_2[0], r),
        G1.mul(pk.g_gamma2_z2, 2 * r * d1 % p),
        G1.mul(c2_acc, r),
        G1.mul(pk.g_gamma2_z_t[0], d2),
        g_acc,
    ])
    return g_a, g_b, g_c


def gm17_trapdoor_check(pk: GM17PK, r1cs: R1CS, z: Sequence[int], d1: int, d2: int, r: int, proof) -> bool:
    """Trapdoor check of a GM17 proof produced with known (d1, d2, r): a = gamma (u + (r + d1) Z) with
    u = sum full_i At_i(t); b = a (in G2); c from the verification equation
    (a + alpha)(a + beta) = alpha beta + gamma psi + c.  Checks A == [a]G, B == [a]H, C == [c]G -- i.e. that
    the proof satisfies both of GM17's pairing equations AND is the honest prover's output for this randomness.
    The SAP quotient is never formed here, so witness map, NTTs and all MSMs are validated end to end."""
    t = pk.trapdoor
    pairing = pk.pairing
    p = pairing.fr.p
    G1, G2 = pairing.g1, pairing.g2
    full = sap_extend_assignment(r1cs, z)
    u = sum(x * y for x, y in zip(full, t["At"])) % p
    a_log = t["gamma"] * ((u + (r + d1) * t["zt"]) % p) % p
    psi = sum(z[i] * t["vk_sc"][i] for i in range(r1cs.num_inputs)) % p
    c_log = ((a_log + t["alpha"]) * (a_log + t["beta"]) - t["alpha"] * t["beta"] - t["gamma"] * psi) % p
    A, B, C = proof
    return (A == G1.mul(generator(G1), a_log) and B == G2.mul(generator(G2), a_log)
            and C == G1.mul(generator(G1), c_log))


# ----------------------------------------------------------------------------------------------
# Self-check of every constant ("verify every recalled constant arithmetically").
# ----------------------------------------------------------------------------------------------
def _is_probable_prime(n: int) -> bool:
    if n < 2:
        return False
    small = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37]
    for q in small:
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def self_check() -> None:
    for fp in (FR4, FQ4):
        assert _is_probable_prime(fp.p) and fp.p.bit_length() == MODULUS_BITS
        assert (fp.p - 1) % (1 << fp.two_adicity) == 0 and (fp.p - 1) % (1 << (fp.two_adicity + 1)) != 0
        w = fp.two_adic_root
        assert pow(w, 1 << fp.two_adicity, fp.p) == 1 and pow(w, 1 << (fp.two_adicity - 1), fp.p) == fp.p - 1
        assert (fp.p * fp.inv64 + 1) % (1 << 64) == 0 and (fp.p * fp.inv32 + 1) % (1 << 32) == 0
    assert FR4.inv64 == 0xBB4334A3FFFFFFFF and FQ4.inv64 == 0xB071A1B67165FFFF
    assert FR4.two_adic_root == 120638817826913173458768829485690099845377008030891618010109772937363554409782252579816313
    assert FQ4.two_adic_root == 264706250571800080758069302369654305530125675521263976034054878017580902343339784464690243
    assert (Q4 - 1) % 49 == 0 and (Q4 - 1) % 343 != 0
    # non-residues
    assert pow(FQ2_NONRESIDUE, (Q4 - 1) // 2, Q4) == Q4 - 1
    assert pow(FQ3_NONRESIDUE, (R4 - 1) // 3, R4) != 1
    # the cycle: embedding degrees 4 and 6
    assert pow(Q4, 4, R4) == 1 and pow(Q4, 2, R4) != 1
    assert pow(R4, 6, Q4) == 1 and pow(R4, 3, Q4) != 1 and pow(R4, 2, Q4) != 1
    for cv in (MNT4_G1, MNT4_G2, MNT6_G1, MNT6_G2):
        g = generator(cv)
        assert cv.is_on_curve(g) and cv.mul(g, cv.order) is None
        P = _find_point(cv, 5)
        assert cv.mul(P, cv.order * cv.cofactor) is None


if __name__ == "__main__":
    self_check()
    print("oracle self-check OK")
