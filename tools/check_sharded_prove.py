#!/usr/bin/env python3
"""Multi-GPU check (torchrun, one rank per GPU; also runs with one process): pcd_b200.sharding.ShardedGroth16 -- MSM
point ranges per GPU, one all_gather of partial sums, assembly on rank 0 -- gives the same proof bytes as the
single-GPU prover and as the instance's known discrete logarithms.  One JSON line on rank 0; exit code 1 on mismatch."""
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
    cases = [(0, int(os.environ.get("LOG_N", "18"))), (1, int(os.environ.get("LOG_N_HELP", "14")))]
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = pcd_b200.Context(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ok_all, out = True, {}
    for pairing, lg in cases:
        inst = synthetic.make_groth16_instance(ctx, pairing, lg, seed=55 + pairing)  # the same instance on every rank
        pk = pcd_b200.ProvingKey(pairing=pairing, **inst["pk"])
        cm = pcd_b200.ConstraintMatrices(pairing, inst["num_inputs"], inst["num_witness"], inst["A"], inst["B"], inst["C"])
        sh = sharding.ShardedGroth16(ctx, pk, cm, rank, world, dev)
        z = torch.from_numpy(inst["z"].view(np.int64)).to(dev)
        p = inst["p"]
        r, s = pow(3, 111, p), pow(7, 99, p)
        for _ in range(2):
            proof = sh.prove(z, r, s)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            proof = sh.prove(z, r, s)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        sh.close()
        if rank == 0:
            g = pcd_b200.Groth16(ctx, pairing)
            idx = g.index(pk, cm, precompute=True)
            lim = lambda v: np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(5)], dtype=np.uint64)
            ref = g.create_proof_dev(idx, z.data_ptr(), lim(r), lim(s)).affine_limbs()
            for _ in range(2):
                g.create_proof_dev(idx, z.data_ptr(), lim(r), lim(s))
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(reps):
                g.create_proof_dev(idx, z.data_ptr(), lim(r), lim(s))
            a1.record(stream)
            torch.cuda.synchronize()
            same = bool(np.array_equal(proof, ref))
            logs_ok = bool(np.array_equal(proof, synthetic.expected_proof(ctx, inst, r, s)))
            out["pairing%d_2^%d" % (pairing, lg)] = {"gpus": world, "matches_single_gpu": same, "matches_trapdoor": logs_ok,
                                                      "ms_sharded": ms, "ms_one_gpu": a0.elapsed_time(a1) / reps}
            ok_all = ok_all and same and logs_ok
            idx.close()
        if world > 1:
            dist.barrier()
    # one MSM sharded by point range through the library's collective (pcdgpu_msm_bases_sharded_dev)
    lg = int(os.environ.get("LOG_N_MSM", "18"))
    n = 1 << lg
    pts = synthetic.random_points_dev(ctx, 0, n, seed=3).cpu().numpy().view(np.uint64)  # same points on every rank
    sc_host = synthetic.random_limbs(n, 0, 9)
    lo, hi = sharding.shard_range(n, world, rank)
    shard = pcd_b200.Bases(ctx, 0, pts[lo:hi], precompute=True)
    sc = torch.from_numpy(sc_host[lo:hi].copy().view(np.int64)).to(dev)
    res = torch.zeros(16, dtype=torch.int64, device=dev)
    for _ in range(3):
        shard.msm_sharded_dev(sc.data_ptr(), hi - lo, res.data_ptr())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        shard.msm_sharded_dev(sc.data_ptr(), hi - lo, res.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms_sh = e0.elapsed_time(e1) / 10
    got = res.cpu().numpy().view(np.uint64)[:10]
    shard.close()
    if rank == 0:
        full = pcd_b200.Bases(ctx, 0, pts, precompute=True)
        ref = full.msm(sc_host)
        scf = torch.from_numpy(sc_host.view(np.int64)).to(dev)
        for _ in range(3):
            full.msm_dev(scf.data_ptr(), n, res.data_ptr())
        e0.record(stream)
        for _ in range(10):
            full.msm_dev(scf.data_ptr(), n, res.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        same = bool(np.array_equal(got, ref))
        out["g1_msm_2^%d" % lg] = {"gpus": world, "matches_single_gpu": same, "ms_sharded": ms_sh,
                                   "ms_one_gpu": e0.elapsed_time(e1) / 10}
        ok_all = ok_all and same
        full.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)
        sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
