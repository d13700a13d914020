#!/usr/bin/env python3
"""Development aid: where does a PCD step's time go?  Prints, for the main (MNT4-298, 2^18), helper (MNT6-298, 2^16)
and tiny default-circuit proofs (2^9 / 2^10, fresh key every call, no tables), the wall / device time per proof and the
CUDA-event timeline of one proof's phases."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import synthetic  # noqa: E402

NAMES = ["sort", "acc_g1", "acc_g2", "reduce", "horner", "ntt", "spmv", "assemble", "acc_g2q3", "acc_small", "acc_tail"]
ctx = pcd_b200.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev, priority=-1 if not os.environ.get("PCDGPU_NO_PRIORITIES") else 0)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
out = {}


def timeline(g, idx, z, r, s, label):
    ctx.lib.pcdgpu_profile_enable(ctx.h, 1)
    g.create_proof_dev(idx, z.data_ptr(), r, s)
    cap = 512
    t0 = (ctypes.c_double * cap)()
    t1 = (ctypes.c_double * cap)()
    cls = (ctypes.c_int * cap)()
    n = ctypes.c_size_t()
    ctx._check(ctx.lib.pcdgpu_profile_timeline(ctx.h, t0, t1, cls, cap, ctypes.byref(n)))
    ctx.lib.pcdgpu_profile_enable(ctx.h, 0)
    rows = sorted((t0[i], t1[i], NAMES[cls[i]]) for i in range(n.value))
    print("---- timeline %s" % label)
    for a, b, nm in rows:
        print("%8.3f -> %8.3f  (%6.3f ms)  %s" % (a, b, b - a, nm))
    sys.stdout.flush()


def one(pairing, log_n, precompute, label, reps=5):
    inst = synthetic.make_groth16_instance(ctx, pairing, log_n,
                                            seed=77 + pairing + (10 * log_n if os.environ.get("PROBE_BENCH_SEED") else 0))
    g = pcd_b200.Groth16(ctx, pairing)
    pk = pcd_b200.ProvingKey(pairing=pairing, **inst["pk"])
    cm = pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"])
    t0 = time.perf_counter()
    if os.environ.get("PCD_MSM_WINDOW"):
        ctx.set_msm_window(int(os.environ["PCD_MSM_WINDOW"]))
    idx = g.index(pk, cm, precompute=precompute)
    ctx.set_msm_window(0)
    ctx.sync()
    t_index = 1e3 * (time.perf_counter() - t0)
    z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
    p = inst["p"]
    r_i, s_i = 0x1234567 * 3 ** 70 % p, 0x7654321 * 5 ** 60 % p
    lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
    proof = g.create_proof_dev(idx, z.data_ptr(), lim(r_i), lim(s_i))
    ok = bool(np.array_equal(proof.affine_limbs(), synthetic.expected_proof(ctx, inst, r_i, s_i)))
    for _ in range(3):
        g.create_proof_dev(idx, z.data_ptr(), lim(r_i), lim(s_i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        g.create_proof_dev(idx, z.data_ptr(), lim(r_i), lim(s_i))
    e1.record(stream)
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - w0) / reps
    ms = e0.elapsed_time(e1) / reps
    # fresh key per call (what the tiny default-circuit proves see): index + prove + free
    w0 = time.perf_counter()
    for _ in range(3):
        i2 = g.index(pk, cm, precompute=precompute)
        g.create_proof_dev(i2, z.data_ptr(), lim(r_i), lim(s_i))
        i2.close()
    fresh = 1e3 * (time.perf_counter() - w0) / 3
    print("== %s: pairing %d domain 2^%d precompute=%s ok=%s  index %.2f ms  prove %.3f ms (wall %.3f)  fresh-key prove %.3f ms"
          % (label, pairing, log_n, precompute, ok, t_index, ms, wall, fresh))
    out[label] = {"pairing": pairing, "log_n": log_n, "precompute": precompute, "ok": ok, "index_ms": t_index,
                  "prove_ms": ms, "wall_ms": wall, "fresh_key_prove_ms": fresh}
    timeline(g, idx, z, lim(r_i), lim(s_i), label)
    idx.close()


which = os.environ.get("PROBE", "main,help,tiny").split(",")
if "main" in which:
    one(0, int(os.environ.get("MAIN_LOG", "18")), True, "main")
if "help" in which:
    one(1, int(os.environ.get("HELP_LOG", "16")), True, "helper")
if "tiny" in which:
    for pairing in (0, 1):
        for lg in (9, 10):
            one(pairing, lg, False, "tiny_p%d_2^%d" % (pairing, lg))
            one(pairing, lg, True, "tinypre_p%d_2^%d" % (pairing, lg))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", os.environ.get("PROBE_OUT", "probe_pcd.json")), "w") as f:
    json.dump(out, f, indent=1)
