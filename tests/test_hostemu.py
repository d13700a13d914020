"""The DEVICE arithmetic headers (pcd_b200/csrc/{fp,fpx,ec}.cuh) compiled for the CPU with an
emulation of the PTX carry-flag primitives (tests/hostemu), checked against the golden vectors and
the C++ oracle.  Catches carry-chain / formula bugs without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import c_oracle as co
import codec
import synth

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "hostemu", "_build", "libhostemu.so")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "hostemu", "hostemu.cpp")
    deps = [src, os.path.join(HERE, "hostemu", "hostemu_prims.h")] + [
        os.path.join(ROOT, "pcd_b200", "csrc", f) for f in ("fp.cuh", "fpx.cuh", "ec.cuh", "prims.cuh", "constants.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I",
                               os.path.join(ROOT, "pcd_b200", "csrc"), "-I", os.path.join(HERE, "hostemu"), src,
                               "-o", SO])
    return ctypes.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


OPS = {"add": 0, "sub": 1, "mul": 2, "sqr": 3, "inv": 4, "neg": 5, "dbl": 10}


def test_fields_golden(emu):
    g = codec.load("fields")
    for field in (0, 1, 2, 3):
        for case in g[str(field)]:
            a, b = codec.hex_to_u64(case["a"]), codec.hex_to_u64(case["b"])
            for name, op in OPS.items():
                out = np.zeros_like(a)
                if field < 2:
                    emu.emu_fp_op(field, op, _p(a), _p(b), _p(out))
                else:
                    emu.emu_fx_op(field, op, _p(a), _p(b), _p(out))
                assert codec.u64_to_hex(out) == case[name], (field, name)


def test_field_random_vs_oracle(emu):
    for field in (0, 1):
        xs = codec.random_field_elems(200, field, 9)
        ys = codec.random_field_elems(200, field, 10)
        for i in range(200):
            for op in (0, 1, 2):
                out = np.zeros(5, dtype=np.uint64)
                emu.emu_fp_op(field, op, _p(xs[i]), _p(ys[i]), _p(out))
                assert np.array_equal(out, co.field_op(field, op, xs[i], ys[i]))


def test_product_of_unreduced_operands(emu):
    """The lane-cooperative group law (wec.cuh) feeds the Montgomery product operands that were never reduced:
    a + k1 p and b + k2 p with (k1 + 1)(k2 + 1) <= 2^20.  The product must still be the canonical a b / R mod p: the
    pre-subtraction value is < p (1 + 2^20 p / 2^320) < 2 p, and no row accumulator leaves its ten limbs."""
    for field in (0, 1):
        p = codec.FIELD_P[field]
        xs = codec.random_field_elems(60, field, 41)
        ys = codec.random_field_elems(60, field, 42)
        mults = [(0, 0), (1, 0), (1, 1), (1023, 1023), ((1 << 20) - 1, 0), (0, (1 << 20) - 1), (2047, 511), (292, 292),
                 (54, 2515)]
        for i in range(60):
            xi, yi = codec.limbs_to_int(xs[i]), codec.limbs_to_int(ys[i])
            want = co.field_op(field, 2, xs[i], ys[i])
            for k1, k2 in mults:
                a = codec.int_to_limbs(xi + k1 * p)
                b = codec.int_to_limbs(yi + k2 * p)
                out = np.zeros(5, dtype=np.uint64)
                emu.emu_fp_op(field, 2, _p(a), _p(b), _p(out))
                assert np.array_equal(out, want), (field, i, k1, k2)
        # extreme operands: the largest values below 2^10 p
        big = codec.int_to_limbs(1024 * p - 1)
        out = np.zeros(5, dtype=np.uint64)
        emu.emu_fp_op(field, 2, _p(big), _p(big), _p(out))
        assert np.array_equal(out, co.field_op(field, 2, codec.int_to_limbs(p - 1), codec.int_to_limbs(p - 1)))


def test_inverse_binary_gcd(emu):
    """Fp::inverse (binary extended Euclid on the Montgomery representative) == Fermat's a^(p-2) == the oracle, on
    edge values (0, 1, 2, p - 1, powers of two, all-ones limbs) and random elements."""
    for field in (0, 1):
        p = codec.FIELD_P[field]
        special = [0, 1, 2, 3, p - 1, p - 2, (1 << 297) - 1, ((1 << 298) - 1) % p, p >> 1, 0xFFFFFFFF, (1 << 288) - 1,
                   1 << 64, 1 << 255, (p + 1) // 2]
        sp = np.stack([codec.int_to_limbs(v % p) for v in special])
        xs = np.concatenate([sp, codec.random_field_elems(300, field, 19)])
        R = (1 << 320) % p
        for i in range(len(xs)):
            out, ref = np.zeros(5, dtype=np.uint64), np.zeros(5, dtype=np.uint64)
            emu.emu_fp_op(field, 4, _p(xs[i]), _p(xs[i]), _p(out))
            emu.emu_fp_op(field, 12, _p(xs[i]), _p(xs[i]), _p(ref))
            assert np.array_equal(out, ref), (field, i)
            a = codec.limbs_to_int(xs[i]) * pow(R, -1, p) % p  # the element this Montgomery representative stands for
            want = pow(a, -1, p) * R % p if a else 0
            assert codec.limbs_to_int(out) == want, (field, i)


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_curve_ops(emu, curve):
    pts = synth.random_points(4, curve, 70 + curve, threads=1)
    P, Q = pts[0], pts[1]
    L = codec.POINT_LIMBS[curve]
    k = codec.random_field_elems(1, codec.SCALAR_FIELD[curve], 5)[0]
    kz = np.zeros(5, dtype=np.uint64)

    def run(op, p, q, kk=kz):
        out = np.zeros(L, dtype=np.uint64)
        emu.emu_ec_op(curve, op, _p(p), _p(q), _p(kk), 10, _p(out))
        return out

    two = codec.int_to_limbs(2)
    P2 = co.fixed_base_mul(curve, P, two.reshape(1, 5), 1)[0]
    Q2 = co.fixed_base_mul(curve, Q, two.reshape(1, 5), 1)[0]
    assert np.array_equal(run(0, P, Q), co.point_sum(curve, np.stack([P, Q])))
    assert np.array_equal(run(0, P, P), P2)                      # madd with equal points -> doubling
    negP = P.copy()
    inf = np.zeros(L, dtype=np.uint64)
    assert np.array_equal(run(0, P, inf), P)
    assert np.array_equal(run(0, inf, Q), Q)
    assert np.array_equal(run(1, P, Q), co.point_sum(curve, np.stack([P2, Q])))
    assert np.array_equal(run(2, P, Q), co.point_sum(curve, np.stack([P2, Q2])))
    assert np.array_equal(run(3, P, Q, k), co.fixed_base_mul(curve, P, k.reshape(1, 5), 1)[0])
    assert np.array_equal(run(4, P, Q), P2)
    assert np.array_equal(run(5, P, Q), co.fixed_base_mul(curve, P, codec.int_to_limbs(4).reshape(1, 5), 1)[0])
    assert not run(6, P, Q).any()                                   # 2P - 2P = infinity
    assert np.array_equal(run(7, P, Q), co.fixed_base_mul(curve, P, codec.int_to_limbs(4).reshape(1, 5), 1)[0])
    del negP
