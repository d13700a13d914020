#!/usr/bin/env python3
"""Workload for `ncu --set full` captures at the PCD step's sizes: the main proof (MNT4-298, 2^18, witness-like
assignment, resident window tables) proved twice with the MSM lanes serialised (one kernel at a time, as ncu replays
them anyway); NCU_STEP=help does the helper proof (MNT6-298, 2^16) instead.  Skip the first proof's launches with
--launch-skip (it grows the scratch)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402

which = os.environ.get("NCU_STEP", "main")
pairing, lg = (0, 18) if which == "main" else (1, 16)
ctx = pcd_b200.Context(0)
dev = torch.device("cuda:0")
inst = synthetic.make_groth16_instance(ctx, pairing, lg, seed=77 + pairing + 10 * lg)
g = pcd_b200.Groth16(ctx, pairing)
idx = g.index(pcd_b200.ProvingKey(pairing=pairing, **inst["pk"]),
              pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"]),
              precompute=True)
z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
p = inst["p"]
lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
r, s = lim(0x1234567 * 3 ** 70 % p), lim(0x7654321 * 5 ** 60 % p)
ctx.set_concurrency(False)
for _ in range(2):
    g.create_proof_dev(idx, z.data_ptr(), r, s)
ctx.sync()
print("ncu step target done:", which)
