#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun, one rank per GPU, N >= 2): the witness map with its a / b / c chains on
different GPUs (pcd_b200.sharding.witness_map_by_vector over pcdgpu_qap_vector_dev / pcdgpu_qap_combine_dev, NCCL
send / recv of the coset evaluations) equals the single-GPU pcdgpu_witness_map bit for bit.  Prints one JSON line
on rank 0; exit code 1 on mismatch."""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcd_b200  # noqa: E402
from pcd_b200 import sharding, synthetic  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    log_n = int(os.environ.get("LOG_N", "18"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pcd_b200.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ok_all = True
    out = {}
    for pairing, lg in ((0, log_n), (1, min(log_n, 16))):
        inst = synthetic.make_groth16_instance(ctx, pairing, lg, seed=99 + pairing)  # same instance on every rank
        g = pcd_b200.Groth16(ctx, pairing)
        idx = g.index(pcd_b200.ProvingKey(pairing=pairing, **inst["pk"]),
                      pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"],
                                                  inst["C"]))
        n = idx.domain_size
        z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
        vp = ctypes.c_void_p

        def vector_fn(which):
            t = torch.empty((n, 5), dtype=torch.int64, device=dev)
            ctx._check(ctx.lib.pcdgpu_qap_vector_dev(ctx.h, idx.r1cs, which, vp(z.data_ptr()), vp(t.data_ptr())))
            return t

        def combine_fn(a, b, c):
            ctx._check(ctx.lib.pcdgpu_qap_combine_dev(ctx.h, idx.r1cs, vp(a.data_ptr()), vp(b.data_ptr()), vp(c.data_ptr())))
            return a

        alloc = lambda: torch.empty((n, 5), dtype=torch.int64, device=dev)
        for _ in range(2):
            h = sharding.witness_map_by_vector(vector_fn, combine_fn, alloc)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            h = sharding.witness_map_by_vector(vector_fn, combine_fn, alloc)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if rank == 0:
            ref = g.witness_map(idx, inst["z"])
            same = bool(np.array_equal(h.cpu().numpy().view(np.uint64), ref))
            # single-GPU time of the same map (device resident), for comparison
            hh = np.zeros((n, 5), dtype=np.uint64)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(reps):
                va, vb, vc = vector_fn(0), vector_fn(1), vector_fn(2)
                combine_fn(va, vb, vc)
            a1.record(stream)
            torch.cuda.synchronize()
            out["pairing%d" % pairing] = {"domain": n, "matches_single_gpu": same, "ms_by_vector": ms,
                                          "ms_one_gpu": a0.elapsed_time(a1) / reps, "gpus": world}
            ok_all = ok_all and same
            del hh
        idx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)
        sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
