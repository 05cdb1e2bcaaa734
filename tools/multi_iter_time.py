"""Per-iteration device time of the partitioned CG under different data-plane settings, one assemble per run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/multi_iter_time.py [workload] [iterations] [ENV=VAL,ENV=VAL ...]

Each extra argument is one setting (comma-separated environment assignments, "-" = defaults); the library reads
STAN_FUSED_HALO at every solve.  Rank 0 prints one JSON line per setting: ms per iteration (max over ranks),
mean SpMV launch time and achieved GB/s on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver, comm_unique_id  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "beam_10m_g2"
its = int(sys.argv[2]) if len(sys.argv) > 2 else 400
settings = sys.argv[3:] or ["-"]
if name in mesh.WORKLOADS:
    m = mesh.workload(name, tolerance=1e-8)
else:                                                 # "nx,ny,nz[,n_parts]"
    f = [int(v) for v in name.split(",")]
    m = mesh.beam(f[0], f[1], f[2], n_parts=f[3] if len(f) > 3 else 1, tolerance=1e-8)
s = Solver(device=local, rank=rank, world=world)
if world > 1:
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(uid[0])
s.SetModel(m); s.AssignDOF(); s.ParallelAssembly_K()
for setting in settings:
    env = dict(kv.split("=") for kv in setting.split(",")) if setting != "-" else {}
    os.environ.update(env)
    rows = []
    for timek in (0, 0, 1):
        if world > 1:
            dist.barrier()
        cg = s.LinearSolver_CG(merit_check=0, IterMax=its, time_kernels=timek)
        t = torch.tensor([cg.solve_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows.append((float(t[0]) / cg.iterationscount, cg))
    for k in env:
        os.environ.pop(k)
    cg = rows[-1][1]
    sp = cg.spmv_ms / max(cg.spmv_launches, 1)
    if rank == 0:
        print(json.dumps({"setting": setting, "world": world, "workload": name, "iterations": cg.iterationscount,
                          "ms_per_iteration": round(rows[1][0], 4), "ms_per_iteration_with_launch_events": round(rows[2][0], 4),
                          "spmv_ms": round(sp, 4), "spmv_gbs": round(cg.spmv_bytes / sp / 1e6, 1), "launches": cg.kernel_launches}), flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()
