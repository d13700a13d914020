#!/usr/bin/env python3
"""Short workload for ncu captures of the lane-cooperative tail kernels (wec.cuh): one MSM per curve over resident
precomputed bases at PCD sizes (G1 2^18, MNT4 G2 2^18, MNT6 G2 2^16)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402

ctx = pcd_b200.Context(0)
dev = torch.device("cuda:0")
res = torch.zeros(64, dtype=torch.int64, device=dev)
for curve, lg in ((3, 16), (1, 18), (0, 18)):
    if os.environ.get("NCU_CURVES") and str(curve) not in os.environ["NCU_CURVES"].split(","):
        continue
    n = 1 << lg
    field = 0 if curve < 2 else 1
    pts = synthetic.random_points_dev(ctx, curve, n, seed=3 + curve)
    sc = torch.from_numpy(synthetic.random_limbs(n, field, 9).view(np.int64)).to(dev)
    ctx.lib.pcdgpu_set_msm_side_by_side(ctx.h, 1)
    bases = pcd_b200.Bases(ctx, curve, pts.cpu().numpy().view(np.uint64), precompute=True)
    for _ in range(2):
        bases.msm_dev(sc.data_ptr(), n, res.data_ptr())
    ctx.sync()
    bases.close()
print("ncu target done")
