import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import c_oracle as co, codec, synth
lib = ctypes.CDLL(os.path.join(ROOT, "tests/cuda/_build/libecprobe.so"))
_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
for curve in (1, 3, 0):
    pts = synth.random_points(2, curve, 70 + curve, threads=1)
    P, Q = pts[0], pts[1]
    L = codec.POINT_LIMBS[curve]
    times = lambda pt, d: co.fixed_base_mul(curve, pt, codec.int_to_limbs(d).reshape(1, 5), 1)[0]
    P2 = times(P, 2)
    want = {0: co.point_sum(curve, np.stack([P, Q])), 1: co.point_sum(curve, np.stack([P2, Q])), 4: P2, 5: times(P, 4),
            7: times(P, 4), 10: co.point_sum(curve, np.stack([P2, Q])), 11: co.point_sum(curve, np.stack([P2, Q])),
            12: co.point_sum(curve, np.stack([P2, Q])), 13: co.point_sum(curve, np.stack([P2, Q])),
            14: co.point_sum(curve, np.stack([P2, Q]))}
    res = {}
    for op, w in want.items():
        out = np.zeros(L, dtype=np.uint64)
        rc = lib.probe_ec_op(curve, op, _p(P), _p(Q), _p(np.zeros(5, dtype=np.uint64)), 10, _p(out))
        res[op] = (rc, bool(np.array_equal(out, w)))
    for d in (3, 5):
        for op in (8, 9):
            out = np.zeros(L, dtype=np.uint64)
            rc = lib.probe_ec_op(curve, op, _p(P), _p(Q), _p(np.zeros(5, dtype=np.uint64)), d, _p(out))
            res[(op, d)] = (rc, bool(np.array_equal(out, times(P, d))))
    print("curve", curve, res)
