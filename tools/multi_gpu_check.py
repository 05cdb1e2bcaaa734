"""Multi-GPU parity: the partitioned solve (one process per GPU, NCCL halo + all-reduce) against the
single-GPU solve of the same model on the same device.  Launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [nx ny nz]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver, comm_unique_id  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims = [int(v) for v in sys.argv[1:4]] if len(sys.argv) >= 4 else [12, 10, 60]
m = mesh.beam(*dims, jitter=True, n_parts=2, tolerance=1e-9)
uid = [comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)

multi = Solver(device=local, rank=rank, world=world)
multi.comm_init(uid[0])
rm = multi.SolverLinearStatics(m, merit_check=0)
from stan_b200 import partition  # noqa: E402
wb = partition.weighted_bounds(m.conn, rm.node_index, world)
assert multi.partition() == (int(wb[rank]), int(wb[rank + 1])), (multi.partition(), wb)   # host mirror of partition_rows
single = Solver(device=local)
rs = single.SolverLinearStatics(m, node_index=rm.node_index, merit_check=0)

du = np.linalg.norm(rm.U_full - rs.U_full) / np.linalg.norm(rs.U_full)
e0, e1 = multi.element_range()
assert rm.stress.shape[0] == e1 - e0 and rs.stress.shape[0] == m.n_elem
ds = np.abs(rm.stress - rs.stress[e0:e1]).max() / np.abs(rs.stress).max()
# post-processing scalars: cell data of the rank's element slice, point data of its rows (DOF-map order)
cell_m, point_m, _ = multi.Load_Scalar()
cell_s, point_s, _ = single.Load_Scalar()
r0, r1 = multi.partition()
inv = np.empty(m.n_nodes, dtype=np.int64)
inv[rm.node_index] = np.arange(m.n_nodes)                    # node at each row
scale_c = np.abs(cell_s).max(axis=(0, 2), keepdims=True) + 1e-30
scale_p = np.abs(point_s).max(axis=0, keepdims=True) + 1e-30
dc = float((np.abs(cell_m - cell_s[e0:e1]) / scale_c).max())
dp = float((np.abs(point_m - point_s[inv[r0:r1]]) / scale_p).max())
ok = (du < 1e-9 and ds < 1e-7 and rm.cg.terminationtype == rs.cg.terminationtype == 1
      and abs(rm.cg.iterationscount - rs.cg.iterationscount) <= 20
      and cell_m.shape[0] == e1 - e0 and point_m.shape[0] == r1 - r0 and dc < 1e-5 and dp < 1e-5)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
print(f"rank {rank}/{world}: rows {multi.partition()}, its {rm.cg.iterationscount} vs {rs.cg.iterationscount}, "
      f"|dU|/|U| = {du:.2e}, |dS|/|S| = {ds:.2e}, scalars cell {dc:.1e} point {dp:.1e}, solve {rm.cg.solve_ms:.1f} ms vs {rs.cg.solve_ms:.1f} ms", flush=True)
multi.close(); single.close()
dist.destroy_process_group()
sys.exit(0 if int(flag[0]) == 1 else 1)
