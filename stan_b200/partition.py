"""Host mirror of the multi-GPU row partition (stan_b200/csrc/comm.cu).

Rank r owns a contiguous range of BFS-ordered nodes; the ranges carry (approximately) equal numbers of stored
3x3 blocks, estimated from the valence of every node (`weighted_bounds`, what `partition_rows` computes on the
device); `bounds` is the plain equal-node-count split the other helpers accept as well.  Because the node
adjacency is symmetric, the set of rows rank r must send to rank s equals the set of halo columns
rank s finds in its own block rows, in the same ascending order — so no set-up communication is
needed.  These helpers restate that logic in numpy for CPU tests and for sizing estimates.
"""
from __future__ import annotations

import numpy as np


def bounds(n_nodes: int, world: int) -> np.ndarray:
    return np.array([n_nodes * r // world for r in range(world + 1)], dtype=np.int64)


def weighted_bounds(conn: np.ndarray, node_index: np.ndarray, world: int) -> np.ndarray:
    """Row bounds balanced by estimated stored blocks: weight 3 * valence + 3 per row (8 incident hexahedra ->
    27 blocks), bound r = one past the first row whose inclusive prefix weight reaches r/W of the total —
    integer arithmetic identical to k_weight_bounds in comm.cu."""
    n = int(node_index.size)
    val = np.bincount(node_index[conn].ravel(), minlength=n).astype(np.int64)
    prefix = np.cumsum(3 * val + 3)
    total = int(prefix[-1])
    b = np.zeros(world + 1, dtype=np.int64)
    b[world] = n
    for r in range(1, world):
        target = (total // world) * r + (total % world) * r // world
        lo = int(np.searchsorted(prefix, target, side="left"))
        b[r] = min(lo + 1, n)
    return np.maximum.accumulate(b)


def owner(q: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.searchsorted(b, q, side="right") - 1


def node_adjacency(conn: np.ndarray, node_index: np.ndarray):
    """Pairs (p, q) of BFS nodes that share an element, as a boolean CSR-like set (small meshes)."""
    bfs = node_index[conn]                                  # (n_elem, 8)
    p = np.repeat(bfs, 8, axis=1).ravel()
    q = np.tile(bfs, (1, 8)).ravel()
    return np.unique(np.stack([p, q], axis=1), axis=0)


def halo_and_send_lists(conn, node_index, world: int):
    """For every rank: halo[r][s] = ascending BFS nodes owned by s that r's rows reference,
    send[r][s] = ascending BFS nodes owned by r that s's rows reference."""
    n = int(node_index.size)
    b = bounds(n, world)
    pairs = node_adjacency(conn, node_index)
    op, oq = owner(pairs[:, 0], b), owner(pairs[:, 1], b)
    halo = [[None] * world for _ in range(world)]
    send = [[None] * world for _ in range(world)]
    for r in range(world):
        for s in range(world):
            if r == s:
                halo[r][s] = send[r][s] = np.zeros(0, np.int64)
                continue
            m = (op == r) & (oq == s)
            halo[r][s] = np.unique(pairs[m, 1])             # columns of my rows owned by s
            send[r][s] = np.unique(pairs[m, 0])             # my rows that touch s
    return b, halo, send
