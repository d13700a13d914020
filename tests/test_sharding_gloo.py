"""World-size-2 (gloo, CPU) test of the MSM point-range sharding logic: partition, gather order and
the combine step, with the C++ oracle standing in for the per-rank GPU partial sums."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_exactly():
    from pcd_b200.sharding import shard_ranges
    for n in (0, 1, 7, 8, 1000, (1 << 20) - 1):
        for world in (1, 2, 3, 4, 8):
            rs = shard_ranges(n, world)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import c_oracle as co
    import synth
    from pcd_b200.sharding import sharded_msm
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        curve = 0
        pts = synth.random_points(n, curve, 91, threads=2)
        sc = synth.random_scalars(n, curve, 92, "W")

        def local_partial(lo, hi):
            # a rank's partial in xyzz form: (x, y, 1, 1) of its affine partial sum (zz = 0 encodes infinity)
            aff = co.msm(curve, pts[lo:hi], sc[lo:hi], threads=2) if hi > lo else np.zeros(10, np.uint64)
            one = co.to_mont(1, np.array([[1, 0, 0, 0, 0]], dtype=np.uint64))[0]
            if not aff.any():
                return np.zeros(20, dtype=np.uint64)
            return np.concatenate([aff, one, one])

        def combine(parts):
            assert parts.shape == (world, 20)
            return co.point_sum(curve, np.stack([p[:10] for p in parts if p[10:15].any()]) if parts.any()
                                else np.zeros((0, 10), np.uint64))

        got = sharded_msm(local_partial, combine, n)
        ref = co.msm(curve, pts, sc, threads=2)
        ret[rank] = bool(np.array_equal(got, ref))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [257, 2])
def test_sharded_msm_world2_gloo(n):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


# ---- witness map by vector (a, b, c chains on different ranks) ----------------------------------------------
def test_vector_owner():
    from pcd_b200.sharding import vector_owner
    assert [vector_owner(k, 1) for k in range(3)] == [0, 0, 0]
    assert [vector_owner(k, 2) for k in range(3)] == [0, 1, 0]
    assert [vector_owner(k, 4) for k in range(3)] == [0, 1, 2]
    assert [vector_owner(k, 8) for k in range(3)] == [0, 1, 2]


def _wm_worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import c_oracle as co
    import codec
    import synth
    from pcd_b200.sharding import witness_map_by_vector
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        pairing, m = 0, 200
        inst = synth.make_instance(pairing, m, seed=31, bitlike=0.3)
        n = co.domain_size(pairing, m + inst["num_inputs"])[0]
        empty = (np.zeros(m + 1, np.uint32), np.zeros(0, np.uint32), np.zeros((0, 5), np.uint64))

        def vector_fn(which):
            # the oracle stands in for pcdgpu_qap_vector_dev: M z (plus the instance rows for A), iFFT, coset FFT
            mats = [inst["A"], inst["B"], inst["C"]]
            z = inst["z"]
            v = np.zeros((n, 5), dtype=np.uint64)
            ptr, col, val = mats[which]
            for i in range(m):
                acc = np.zeros(5, dtype=np.uint64)
                for k in range(int(ptr[i]), int(ptr[i + 1])):
                    acc = co.field_op(pairing, 0, acc, co.field_op(pairing, 2, val[k], z[col[k]]))
                v[i] = acc
            if which == 0:
                for j in range(inst["num_inputs"]):
                    v[m + j] = z[j]
            v = co.ntt(pairing, co.ntt(pairing, v, True, False, 1), False, True, 1)
            return torch.from_numpy(v.view(np.int64).copy())

        def combine_fn(a, b, c):
            a, b, c = (t.numpy().view(np.uint64) for t in (a, b, c))
            p = codec.FIELD_P[pairing]
            g = 10
            zinv = pow((pow(g, n, p) - 1) % p, -1, p)
            zl = co.to_mont(pairing, codec.int_to_limbs(zinv).reshape(1, 5))[0]
            h = np.stack([co.field_op(pairing, 2, co.field_op(pairing, 1, co.field_op(pairing, 2, a[i], b[i]), c[i]), zl)
                          for i in range(n)])
            return co.ntt(pairing, h, True, True, 1)

        h = witness_map_by_vector(vector_fn, combine_fn, lambda: torch.empty((n, 5), dtype=torch.int64))
        if rank == 0:
            ref = co.witness_map(pairing, inst["A"], inst["B"], inst["C"], m, inst["num_inputs"], inst["z"], 2)
            ret[rank] = bool(np.array_equal(h, ref))
        else:
            ret[rank] = h is None
        del empty
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_witness_map_by_vector_gloo(world):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_wm_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_witness_map_by_vector_multi_gpu():
    """real GPUs: needs at least two (skipped on a one-GPU box; the gloo tests above cover the host logic)"""
    import subprocess
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 3 if ngpu >= 3 else 2
    env = dict(os.environ, LOG_N="14")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tools", "check_wm_by_vector.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_sharded_groth16_prove_multi_gpu():
    """one Groth16 proof over the GPUs of the box (MSM point ranges per GPU, gathered partial sums) == the single-GPU
    proof == the instance's discrete logarithms; with one GPU the same code runs as a single rank"""
    import subprocess
    import torch
    ngpu = torch.cuda.device_count()
    world = min(ngpu, 4)
    env = dict(os.environ, LOG_N="14", LOG_N_HELP="12")
    script = os.path.join(ROOT, "tools", "check_sharded_prove.py")
    if world < 2:
        cmd = [sys.executable, script]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
