"""Host logic of the PCD-tree scheduler (pcd_b200/tree.py): children before parents, round-robin ownership, per-node
randomness independent of the number of ranks; and, with two gloo ranks, the same tree as one rank produces."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcd_b200 import tree  # noqa: E402

P = 475922286169261325753349249653048451545124878552823515553267735739164647307408490559963137


def fake_prove(node, child_proofs, rng):
    """stands for ECCyclePCD::prove: depends on the node, on its children's proofs and on its own randomness"""
    r = tree.draw_scalar(rng, P)
    s = tree.draw_scalar(rng, P)
    h = hashlib.sha256()
    h.update(str(node).encode())
    for c in child_proofs:
        h.update(c)
    h.update(r.tobytes())
    h.update(s.tobytes())
    return h.digest()


@pytest.mark.parametrize("n", [1, 2, 3, 7, 10, 64])
def test_rounds_respect_dependencies(n):
    seen = set()
    for rnd in tree.rounds(n):
        for node in rnd:
            assert all(c in seen for c in tree.children(node, n)), node
        seen.update(rnd)
    assert seen == set(range(1, n + 1))


def test_64_node_tree_shape():
    sizes = [len(r) for r in tree.rounds(64)]
    assert sum(sizes) == 64 and sizes[0] == 32 and sizes[-1] == 1
    for world in (1, 2, 4, 8):
        sch = tree.schedule(64, world)
        load = [sum(1 for rnd in sch for _, r in rnd if r == k) for k in range(world)]
        assert sum(load) == 64 and max(load) - min(load) <= len(sch)


def test_draw_scalar_below_p_and_deterministic():
    a = tree.draw_scalar(tree.node_rng(5, 9), P)
    b = tree.draw_scalar(tree.node_rng(5, 9), P)
    c = tree.draw_scalar(tree.node_rng(5, 10), P)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert sum(int(a[i]) << (64 * i) for i in range(5)) < P


def test_tree_independent_of_world_size_single_process():
    ref = tree.prove_tree(13, fake_prove)
    for world in (2, 3, 4):
        # simulate the ranks one after the other with a shared mailbox as the exchange
        done = {}
        sch = tree.schedule(13, world)
        for rnd in sch:
            this = {}
            for node, r in rnd:
                this[node] = fake_prove(node, [done[c] for c in tree.children(node, 13)], tree.node_rng(20261017, node))
            done.update(this)
        assert done == ref


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = tree.prove_tree(10, fake_prove, rank=rank, world=world)
    ret[rank] = out == tree.prove_tree(10, fake_prove)
    dist.destroy_process_group()


def test_tree_two_gloo_ranks():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
