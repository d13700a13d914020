"""PCD trees over the GPUs of a box (BASELINE config 5): a host scheduler above the C ABI.

In the reference a PCD node is proved by `ECCyclePCD::prove(pk, predicate, msg, witness, prior_msgs, prior_proofs, rng)`
(/root/reference/src/ec_cycle_pcd/mod.rs:92-181): the node's circuit takes the proofs of its children as witnesses
(`prior_proofs`, mod.rs:165-166; the circuits' loops over `PRIOR_MSG_LEN` priors, data_structures.rs:171-212), so a
parent can only be proved after its children, while nodes that do not depend on each other -- all the leaves, then all
the nodes whose children are done, ... -- are independent proving jobs.  Inside a node the main proof precedes the
helper proof (mod.rs:171,179).  There is no exchange on the data path: the host hands the children's proofs (a few
hundred bytes) to the rank that proves the parent.

Scheduling: nodes are heap-indexed (node i has children 2i and 2i + 1); round h proves the nodes of height h (leaves
first); inside a round node number k (in increasing index order) goes to rank k mod world.  Every node draws its
randomness from its own generator seeded by (tree seed, node index), so the proofs do not depend on the number of GPUs
or on the order in which a rank serves its nodes -- the reference API takes the caller's `&mut rng`
(/root/reference/src/lib.rs:44-52), and concurrently proved nodes need independent streams.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def children(node: int, n_nodes: int) -> List[int]:
    return [c for c in (2 * node, 2 * node + 1) if c <= n_nodes]


def heights(n_nodes: int) -> Dict[int, int]:
    h = {}
    for node in range(n_nodes, 0, -1):
        ch = children(node, n_nodes)
        h[node] = 1 + max(h[c] for c in ch) if ch else 0
    return h


def rounds(n_nodes: int) -> List[List[int]]:
    """nodes grouped by height, leaves first: every node of a round only depends on nodes of earlier rounds"""
    h = heights(n_nodes)
    out: List[List[int]] = [[] for _ in range(max(h.values()) + 1)] if n_nodes else []
    for node in range(1, n_nodes + 1):
        out[h[node]].append(node)
    return out


def owner(position_in_round: int, world: int) -> int:
    return position_in_round % world


def schedule(n_nodes: int, world: int) -> List[List[Tuple[int, int]]]:
    """per round, the list of (node, rank)"""
    return [[(node, owner(k, world)) for k, node in enumerate(rnd)] for rnd in rounds(n_nodes)]


def node_rng(tree_seed: int, node: int) -> np.random.Generator:
    """the node's own deterministic generator (counter-based: Philox keyed by the tree seed and the node index)"""
    return np.random.Generator(np.random.Philox(key=[int(tree_seed) & (2 ** 64 - 1), int(node)]))


def draw_scalar(rng: np.random.Generator, p: int) -> np.ndarray:
    """one uniform scalar below p as five plain-integer u64 limbs (the shape of ark-ff's sampler, SURVEY.md B.6:
    298-bit draws, rejection)"""
    while True:
        limbs = rng.integers(0, 2 ** 64, 5, dtype=np.uint64)
        limbs[4] &= np.uint64((1 << (298 - 256)) - 1)
        v = sum(int(limbs[i]) << (64 * i) for i in range(5))
        if v < p:
            return limbs


def prove_tree(n_nodes: int, prove_node: Callable[[int, Sequence[bytes], np.random.Generator], bytes],
               rank: int = 0, world: int = 1, tree_seed: int = 20261017,
               exchange: Optional[Callable[[Dict[int, bytes]], Dict[int, bytes]]] = None) -> Dict[int, bytes]:
    """Proves the tree.  prove_node(node, child_proofs, rng) -> the node's proof bytes; it is called on the node's
    rank only, after the proofs of its children have arrived.  exchange(mine) -> the union over ranks of this round's
    {node: proof} (default: torch.distributed.all_gather_object when a process group is initialised, identity
    otherwise).  Returns {node: proof} for the whole tree on every rank."""
    if exchange is None:
        exchange = _default_exchange
    done: Dict[int, bytes] = {}
    for rnd in schedule(n_nodes, world):
        mine: Dict[int, bytes] = {}
        for node, r in rnd:
            if r != rank:
                continue
            ch = children(node, n_nodes)
            missing = [c for c in ch if c not in done]
            if missing:
                raise RuntimeError("node %d scheduled before its children %s" % (node, missing))
            mine[node] = prove_node(node, [done[c] for c in ch], node_rng(tree_seed, node))
        done.update(exchange(mine) if world > 1 else mine)
    return done


def _default_exchange(mine: Dict[int, bytes]) -> Dict[int, bytes]:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return mine
    box: List[Optional[Dict[int, bytes]]] = [None] * dist.get_world_size()
    dist.all_gather_object(box, mine)
    out: Dict[int, bytes] = {}
    for d in box:
        out.update(d or {})
    return out
