#!/usr/bin/env python3
"""Development aid: per-proof device time and event timeline inside bench.py's own Step object (all four keys of the PCD
step resident at once, one context), to compare with tools/probe_pcd.py's one-key-at-a-time numbers."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pcd_b200  # noqa: E402

NAMES = ["sort", "acc_g1", "acc_g2", "reduce", "horner", "ntt", "spmv", "assemble", "acc_g2q3", "acc_small", "acc_tail"]
args = argparse.Namespace(pcd_main_log_n=18, pcd_help_log_n=16, pcd_tiny_log_n=10)
ctx = pcd_b200.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev, priority=-1)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
step = bench.PcdStep(ctx, dev, args)
rng = np.random.Generator(np.random.Philox(1))
rs = step.draw(rng)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(2):
    for part, (r, s) in zip(step.parts, rs):
        for _ in range(2):
            part["g"].create_proof_dev(part["idx"], part["z_dev"].data_ptr(), r, s)
        e0.record(stream)
        for _ in range(3):
            part["g"].create_proof_dev(part["idx"], part["z_dev"].data_ptr(), r, s)
        e1.record(stream)
        torch.cuda.synchronize()
        print("== %s: %.3f ms" % (part["label"], e0.elapsed_time(e1) / 3))
for part, (r, s) in zip(step.parts, rs):
    if part["label"] not in os.environ.get("TIMELINE", "helper").split(","):
        continue
    ctx.lib.pcdgpu_profile_enable(ctx.h, 1)
    part["g"].create_proof_dev(part["idx"], part["z_dev"].data_ptr(), r, s)
    cap = 512
    t0, t1, cls, n = (ctypes.c_double * cap)(), (ctypes.c_double * cap)(), (ctypes.c_int * cap)(), ctypes.c_size_t()
    ctx._check(ctx.lib.pcdgpu_profile_timeline(ctx.h, t0, t1, cls, cap, ctypes.byref(n)))
    ctx.lib.pcdgpu_profile_enable(ctx.h, 0)
    print("---- timeline %s" % part["label"])
    for a, b, nm in sorted((t0[i], t1[i], NAMES[cls[i]]) for i in range(n.value)):
        print("  %8.3f -> %8.3f  (%6.3f ms)  %s" % (a, b, b - a, nm))
