"""The device curve arithmetic (ec.cuh) run on the GPU through tests/cuda/ec_probe.cu and compared
with the C++ oracle: the hardware twin of tests/test_hostemu.py."""
import ctypes
import os

import numpy as np
import pytest

import c_oracle as co
import codec
import synth

pytestmark = pytest.mark.gpu
SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda", "_build", "libecprobe.so")


@pytest.fixture(scope="module")
def probe():
    if not os.path.exists(SO):
        pytest.fail("tests/cuda/_build/libecprobe.so missing: run __graft_entry__.build()")
    lib = ctypes.CDLL(SO)
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_curve_ops_on_gpu(probe, curve):
    pts = synth.random_points(2, curve, 70 + curve, threads=1)
    P, Q = pts[0], pts[1]
    L = codec.POINT_LIMBS[curve]
    k = codec.random_field_elems(1, codec.SCALAR_FIELD[curve], 5)[0]
    kz = np.zeros(5, dtype=np.uint64)

    def run(op, p, q, kk=kz, kl=10):
        out = np.zeros(L, dtype=np.uint64)
        assert probe.probe_ec_op(curve, op, _p(p), _p(q), _p(kk), kl, _p(out)) == 0
        return out

    def times(pt, d):
        return co.fixed_base_mul(curve, pt, codec.int_to_limbs(d).reshape(1, 5), 1)[0]

    inf = np.zeros(L, dtype=np.uint64)
    P2, Q2 = times(P, 2), times(Q, 2)
    assert np.array_equal(run(0, P, Q), co.point_sum(curve, np.stack([P, Q])))
    assert np.array_equal(run(0, P, P), P2)
    assert np.array_equal(run(0, P, inf), P)
    assert np.array_equal(run(0, inf, Q), Q)
    assert np.array_equal(run(1, P, Q), co.point_sum(curve, np.stack([P2, Q])))
    assert np.array_equal(run(2, P, Q), co.point_sum(curve, np.stack([P2, Q2])))
    assert np.array_equal(run(3, P, Q, k), co.fixed_base_mul(curve, P, k.reshape(1, 5), 1)[0])
    assert np.array_equal(run(4, P, Q), P2)
    assert np.array_equal(run(5, P, Q), times(P, 4))
    assert not run(6, P, Q).any()
    assert np.array_equal(run(7, P, Q), times(P, 4))
    for d in (1, 2, 3, 5, 15):
        assert np.array_equal(run(8, P, Q, kl=d), times(P, d)), d
        assert np.array_equal(run(9, P, Q, kl=d), times(P, d)), d
