#!/usr/bin/env python3
"""Development probe: per-phase times of one resident-table G1/G2 MSM for several window sizes."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402

ctx = pcd_b200.Context(0)
if os.environ.get("SIDE_BY_SIDE"):
    ctx.lib.pcdgpu_set_msm_side_by_side(ctx.h, 1)
dev = torch.device("cuda:0")
NAMES = ["sort", "acc_g1", "acc_g2", "reduce", "horner", "ntt", "spmv", "assemble", "acc_g2q3", "acc_small", "acc_tail"]


def profile(fn, reps=3):
    fn()
    ctx.sync()
    ctx.lib.pcdgpu_profile_enable(ctx.h, 1)
    for _ in range(reps):
        fn()
    ms = (ctypes.c_double * 10)()
    units = (ctypes.c_double * 10)()
    spans = (ctypes.c_uint64 * 10)()
    launches = ctypes.c_uint64()
    ctx._check(ctx.lib.pcdgpu_profile_read(ctx.h, ms, units, spans, ctypes.byref(launches)))
    ctx.lib.pcdgpu_profile_enable(ctx.h, 0)
    out = {NAMES[i]: round(ms[i] / reps, 3) for i in range(10) if ms[i] > 0}
    ent = sum(units[i] for i in (1, 2, 8, 9)) / reps
    out['entries'] = int(ent)
    return out


curves = [int(x) for x in os.environ.get("PROBE_CURVES", "0").split(",")]
for curve in curves:
    for log_n in [int(x) for x in os.environ.get("PROBE_LOGN", "16,20").split(",")]:
        n = 1 << log_n
        pts = synthetic.random_points_dev(ctx, curve, n, seed=3).cpu().numpy().view(np.uint64)
        sc_u = synthetic.random_limbs(n, 0, 9)
        sc_w = sc_u.copy()
        rng = np.random.Generator(np.random.Philox(5))
        u = rng.random(n)
        sc_w[u < 0.4] = 0
        sc_w[u < 0.2, 0] = 1
        res = torch.zeros(64, dtype=torch.int64, device=dev)
        for c in [int(x) for x in os.environ.get("PROBE_C", "0").split(",")]:
            ctx.set_msm_window(c)
            b = pcd_b200.Bases(ctx, curve, pts, precompute=True)
            for name, sc in (("U", sc_u), ("W", sc_w)):
                d = torch.from_numpy(sc.view(np.int64)).to(dev)
                p = profile(lambda: b.msm_dev(d.data_ptr(), n, res.data_ptr()))
                acc = sum(v for k, v in p.items() if k.startswith("acc"))
                prod = {0: 10, 1: 28, 2: 10, 3: 58}[curve]
                print("curve %d 2^%d c=%d %s acc %.3f ms = %.2f TIMAD/s (%.2f of 9.15)" % (
                    curve, log_n, c, name, acc, p["entries"] * prod * 210 / (acc * 1e-3) / 1e12,
                    p["entries"] * prod * 210 / (acc * 1e-3) / 9.15e12), p, flush=True)
            b.close()
        ctx.set_msm_window(0)
