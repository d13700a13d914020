#!/usr/bin/env python3
"""Short workload for `ncu --set full` captures: a few G1 MSMs at 2^20 over resident precomputed
bases (uniform scalars), one G2 MSM at 2^18 and coset NTTs at 2^22 (r4)."""
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
n = 1 << int(os.environ.get("NCU_MSM_LOG_N", "20"))
pts = synthetic.random_points_dev(ctx, 0, n, seed=3)
sc = torch.from_numpy(synthetic.random_limbs(n, 0, 9).view(np.int64)).to(dev)
res = torch.zeros(64, dtype=torch.int64, device=dev)
bases = pcd_b200.Bases(ctx, 0, pts.cpu().numpy().view(np.uint64), precompute=True)
for _ in range(4):
    bases.msm_dev(sc.data_ptr(), n, res.data_ptr())
ctx.sync()
n2 = 1 << 18
pts2 = synthetic.random_points_dev(ctx, 1, n2, seed=4)
for _ in range(2):
    ctx.msm_dev(1, pts2.data_ptr(), sc.data_ptr(), n2, res.data_ptr())
ctx.sync()
log_n = 22
x = torch.from_numpy(synthetic.random_limbs(1 << 20, 0, 5).view(np.int64)).to(dev).repeat(4, 1)
for _ in range(3):
    ctx.ntt_dev(0, x.data_ptr(), log_n, False, True)
ctx.sync()
print("ncu target done")
