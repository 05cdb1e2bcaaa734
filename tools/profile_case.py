"""Small driver for ncu captures: one assemble + short CG + recovery on a named workload.

    python tools/profile_case.py [workload] [cg_iters] [spmv_reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "beam_100k_g2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
if name in mesh.WORKLOADS:
    m = mesh.workload(name, tolerance=1e-8)
else:                                                 # "nx,ny,nz"
    f = [int(v) for v in name.split(",")]
    m = mesh.beam(f[0], f[1], f[2], tolerance=1e-8)
with Solver() as s:
    s.SetModel(m)
    s.AssignDOF()
    s.ParallelAssembly_K()            # first call pays pool growth and module load
    a = s.ParallelAssembly_K()
    cg = s.LinearSolver_CG(merit_check=0, IterMax=iters)
    r = s.Recovery_Stress()
    ms, by = s.time_spmv(reps)
    print(f"{name}: assembly {a.assembly_ms:.3f} ms, pattern {a.pattern_ms:.3f} ms, cg {cg.iterationscount} its "
          f"{cg.solve_ms:.3f} ms, recovery {r.recover_ms:.3f} ms, spmv {ms:.4f} ms = {by / ms / 1e6:.1f} GB/s")
