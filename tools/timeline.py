#!/usr/bin/env python3
"""Development aid: timeline (CUDA-event spans) of one Groth16 proof with the MSM lanes overlapping."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402

log_n = int(os.environ.get("LOG_N", "20"))
ctx = pcd_b200.Context(0)
dev = torch.device("cuda:0")
PAIRING = int(os.environ.get("PAIRING", "0"))
inst = synthetic.make_groth16_instance(ctx, PAIRING, log_n)
g = pcd_b200.Groth16(ctx, PAIRING)
idx = g.index(pcd_b200.ProvingKey(pairing=PAIRING, **inst["pk"]),
              pcd_b200.ConstraintMatrices(PAIRING, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"]),
              precompute=True)
z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
r = np.array([5, 6, 7, 8, 0], dtype=np.uint64)
for _ in range(3):
    g.create_proof_dev(idx, z.data_ptr(), r, r)
ctx.lib.pcdgpu_profile_enable(ctx.h, 1)
g.create_proof_dev(idx, z.data_ptr(), r, r)
cap = 256
t0 = (ctypes.c_double * cap)()
t1 = (ctypes.c_double * cap)()
cls = (ctypes.c_int * cap)()
n = ctypes.c_size_t()
ctx._check(ctx.lib.pcdgpu_profile_timeline(ctx.h, t0, t1, cls, cap, ctypes.byref(n)))
names = ["sort", "acc_g1", "acc_g2", "reduce", "horner", "ntt", "spmv", "assemble", "acc_g2q3", "acc_small", "acc_tail"]
rows = sorted((t0[i], t1[i], names[cls[i]]) for i in range(n.value))
for a, b, nm in rows:
    print("%8.3f -> %8.3f  (%6.3f ms)  %s" % (a, b, b - a, nm))
