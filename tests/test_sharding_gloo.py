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
