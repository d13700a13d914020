"""The lane schedules of the warp-cooperative group law (tools/gen_wec.py -> pcd_b200/csrc/wec_programs.cuh), run with
Python integers and compared with the oracle's affine group law on every curve: addition (both halves), mixed
addition, doubling and normalisation.  Also checks the hazard rule the GPU interpreter relies on (no slot written by one
lane is touched by another lane inside a step) and that the committed header is what the generator produces."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_wec as gw  # noqa: E402
import pcd_oracle as po  # noqa: E402

CURVES = {0: po.MNT4_G1, 1: po.MNT4_G2, 2: po.MNT6_G1, 3: po.MNT6_G2}


def base_p(curve):
    return po.Q4 if curve in (0, 1) else po.R4


def coeffs(cv, e):
    return list(cv.F.coeffs(e))


def to_xyzz(cv, P, lam):
    """affine -> xyzz with a non-trivial denominator lam (X = x lam^2, Y = y lam^3, ZZ = lam^2, ZZZ = lam^3)"""
    F = cv.F
    l2 = F.sqr(lam)
    l3 = F.mul(l2, lam)
    return [F.mul(P[0], l2), F.mul(P[1], l3), l2, l3]


def flat(cv, elems):
    out = []
    for e in elems:
        out += coeffs(cv, e)
    return out


def from_flat(cv, v, k, n):
    return [cv.F.from_coeffs(v[i * k:(i + 1) * k]) for i in range(n)]


def xyzz_to_affine(cv, v):
    F = cv.F
    x, y, zz, zzz = v
    if F.is_zero(zz):
        return None
    return (F.mul(x, F.inv(zz)), F.mul(y, F.inv(zzz)))


def some_elem(cv, seed):
    k = len(coeffs(cv, cv.F.from_int(1)))
    rng = po.SplitMix64(seed)
    pp = po.Q4 if cv.name.startswith("mnt4") else po.R4
    return cv.F.from_coeffs([rng.field(pp) for _ in range(k)])


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_programs_match_the_oracle(curve):
    cv = CURVES[curve]
    k = gw.CURVES[curve]["k"]
    p = base_p(curve)
    G0 = po.generator(cv)
    P = cv.mul(G0, 0x1234567)
    Q = cv.mul(G0, 0x7654321)
    lam1, lam2 = some_elem(cv, 11 + curve), some_elem(cv, 29 + curve)
    built = {name: gw.build(curve, name) for name in gw.PROGRAMS}
    # addition of two xyzz points
    X = flat(cv, to_xyzz(cv, P, lam1))
    Y = flat(cv, to_xyzz(cv, Q, lam2))
    T = [0] * gw.MAX_T
    gw.interpret(built["add1"], p, X, Y, T)
    Pd = from_flat(cv, T, k, 2)
    assert not cv.F.is_zero(Pd[0])
    gw.interpret(built["add2"], p, X, Y, T)
    assert xyzz_to_affine(cv, from_flat(cv, X, k, 4)) == cv.add(P, Q)
    # add1 detects P == Q and P == -Q through (P, R)
    X = flat(cv, to_xyzz(cv, P, lam1))
    Y = flat(cv, to_xyzz(cv, P, lam2))
    T = [0] * gw.MAX_T
    gw.interpret(built["add1"], p, X, Y, T)
    Pd = from_flat(cv, T, k, 2)
    assert cv.F.is_zero(Pd[0]) and cv.F.is_zero(Pd[1])
    Y = flat(cv, to_xyzz(cv, cv.neg(P), lam2))
    gw.interpret(built["add1"], p, X, Y, T)
    Pd = from_flat(cv, T, k, 2)
    assert cv.F.is_zero(Pd[0]) and not cv.F.is_zero(Pd[1])
    # mixed addition
    X = flat(cv, to_xyzz(cv, P, lam1))
    Y = flat(cv, [Q[0], Q[1]])
    T = [0] * gw.MAX_T
    gw.interpret(built["madd1"], p, X, Y, T)
    gw.interpret(built["madd2"], p, X, Y, T)
    assert xyzz_to_affine(cv, from_flat(cv, X, k, 4)) == cv.add(P, Q)
    # doubling, twice in a row on the same slots
    X = flat(cv, to_xyzz(cv, P, lam1))
    gw.interpret(built["dbl"], p, X, [0] * 16)
    assert xyzz_to_affine(cv, from_flat(cv, X, k, 4)) == cv.double(P)
    gw.interpret(built["dbl"], p, X, [0] * 16)
    assert xyzz_to_affine(cv, from_flat(cv, X, k, 4)) == cv.mul(P, 4)
    # normalisation
    X = flat(cv, to_xyzz(cv, Q, lam2))
    gw.interpret(built["toaff"], p, X, [0] * 16)
    assert tuple(from_flat(cv, X, k, 2)) == tuple(Q)


def test_scalar_multiplication_by_the_programs():
    """double-and-add driven by the schedules (the shape of the Straus / window-combination kernels)"""
    curve = 0
    cv = CURVES[curve]
    p = base_p(curve)
    k = 1
    P = po.generator(cv)
    dbl, a1, a2 = (gw.build(curve, n) for n in ("dbl", "add1", "add2"))
    scalar = 0xdeadbeefcafef00d1234567
    Pq = flat(cv, to_xyzz(cv, P, some_elem(cv, 3)))
    acc = None
    for bit in bin(scalar)[2:]:
        if acc is not None:
            gw.interpret(dbl, p, acc, [0] * 16)
        if bit == "1":
            if acc is None:
                acc = list(Pq)
            else:
                T = [0] * gw.MAX_T
                Y = list(Pq)
                gw.interpret(a1, p, acc, Y, T)
                gw.interpret(a2, p, acc, Y, T)
    assert xyzz_to_affine(cv, from_flat(cv, acc, k, 4)) == cv.mul(P, scalar)


def test_header_is_current():
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_wec.py"), "--check"])
    assert rc == 0, "pcd_b200/csrc/wec_programs.cuh is stale: run python tools/gen_wec.py"
