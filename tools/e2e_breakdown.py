"""Wall-clock breakdown of one end-to-end pass through the C ABI (host buffers in, host buffers out), itemising
what bench.py's `e2e` adds to the device time.  One GPU: `python tools/e2e_breakdown.py [workload] [cg its]`;
several: launch under torchrun (each rank prints its own line, rank 0 last).

STAN_TRACE=1 additionally makes the library print the host wall time of the phases inside stan_assemble."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver, comm_unique_id  # noqa: E402

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
name = sys.argv[1] if len(sys.argv) > 1 else "beam_10m_g2"
its = int(sys.argv[2]) if len(sys.argv) > 2 else 50
m = mesh.workload(name, tolerance=1e-8)
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = Solver(device=local, rank=rank, world=world)
if world > 1:
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    s.comm_init(uid[0])
s.SetModel(m); ni = s.AssignDOF(); s.ParallelAssembly_K(); s.LinearSolver_CG(merit_check=0, IterMax=5); s.Recovery_Stress()
s.Include_BC_DOF(); s.strain_stress()                      # warm: pools, page cache
rows = []
for rep in range(2):
    if world > 1:
        dist.barrier()
    t = [time.perf_counter()]
    def lap():
        t.append(time.perf_counter())
    s.SetModel(m); lap()
    s.SetDOF(ni); lap()
    s.ParallelAssembly_K(); lap()
    s.LinearSolver_CG(merit_check=0, IterMax=its); lap()
    s.Recovery_Stress(); lap()
    U = s.Include_BC_DOF(); lap()
    st = s.strain_stress(); lap()
    disp = s.node_displacements() if hasattr(s, "node_displacements") else U.reshape(-1, 3)[ni]; lap()
    chk = float(np.abs(U).max()); lap()
    names = ["SetModel", "SetDOF", "assemble", f"solve({its} its)", "recover", "get U", "get strain/stress", "disp per node", "max|U|"]
    rows.append({n: round(1e3 * (b - a), 1) for n, a, b in zip(names, t, t[1:])})
    rows[-1]["total_ms"] = round(1e3 * (t[-1] - t[0]), 1)
print(json.dumps({"rank": rank, "world": world, "workload": name, "passes_ms": rows}), flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()
