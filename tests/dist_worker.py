"""world_size-2 gloo worker for tests/test_distributed_cpu.py (launched with torchrun)."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh, partition  # noqa: E402
from oracle import oracle as O  # noqa: E402  (tests may use the oracle)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

# 1. the 128-byte communicator id travels from rank 0 to every rank unchanged
uid = [os.urandom(128) if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
digest = torch.tensor(list(hashlib.sha256(uid[0]).digest()), dtype=torch.int64)
ref = digest.clone()
dist.broadcast(ref, src=0)
assert torch.equal(digest, ref) and len(uid[0]) == 128

# 2. every rank derives the same partition; what r sends to s is what s expects from r
m = mesh.beam(4, 3, 12, jitter=True)
ni = O.assign_dof(m)
b, halo, send = partition.halo_and_send_lists(m.conn, ni, world)
assert b[0] == 0 and b[-1] == m.n_nodes and np.all(np.diff(b) > 0)
mine = [send[rank][s] for s in range(world)]
gathered = [None] * world
dist.all_gather_object(gathered, mine)
for s in range(world):
    if s != rank:
        assert np.array_equal(gathered[s][rank], halo[rank][s]), "send list of the owner != halo list of the reader"
        assert np.all((halo[rank][s] >= b[s]) & (halo[rank][s] < b[s + 1]))

# 3. partitioned SpMV with halo exchange (numpy stand-in for the device kernel) equals the global product
red, _ = O.spc_reduction(m, ni)
K = O.assemble_upper(m, ni, red).to_scipy_full().tocsr()
x = np.random.default_rng(7).standard_normal(K.shape[0])
free = np.where(red != -1)[0]
rows = free[(free // 3 >= b[rank]) & (free // 3 < b[rank + 1])]
loc = np.searchsorted(free, rows)
y_loc = K[loc] @ x                                           # owner computes its rows; x entries of other ranks = halo
parts = [None] * world
dist.all_gather_object(parts, (loc, y_loc))
y = np.zeros_like(x)
for l, v in parts:
    y[l] = v
assert np.allclose(y, K @ x, rtol=1e-13, atol=1e-9)

# 4. dot products: local partial sums + all-reduce == global dot
t = torch.tensor([float(x[loc] @ y_loc)], dtype=torch.float64)
dist.all_reduce(t)
assert abs(t.item() - float(x @ (K @ x))) <= 1e-10 * abs(t.item())
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
