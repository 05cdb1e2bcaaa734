"""Wall-clock breakdown of one end-to-end pass through the C ABI (host buffers in, host buffers out)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver  # noqa: E402

m = mesh.workload(sys.argv[1] if len(sys.argv) > 1 else "beam_10m_g2", tolerance=1e-8)
its = int(sys.argv[2]) if len(sys.argv) > 2 else 50
with Solver() as s:
    s.SetModel(m); ni = s.AssignDOF(); s.ParallelAssembly_K(); s.LinearSolver_CG(merit_check=0, IterMax=5); s.Recovery_Stress()
    s.Include_BC_DOF(); s.strain_stress()                      # warm: pools, page cache
    for rep in range(2):
        t = [time.perf_counter()]
        s.SetModel(m); t.append(time.perf_counter())
        s.SetDOF(ni); t.append(time.perf_counter())
        s.ParallelAssembly_K(); t.append(time.perf_counter())
        s.LinearSolver_CG(merit_check=0, IterMax=its); t.append(time.perf_counter())
        s.Recovery_Stress(); t.append(time.perf_counter())
        U = s.Include_BC_DOF(); t.append(time.perf_counter())
        st = s.strain_stress(); t.append(time.perf_counter())
        names = ["SetModel", "SetDOF", "assemble", f"solve({its} its)", "recover", "get U", "get strain/stress"]
        print(" | ".join(f"{n} {1e3 * (b - a):.0f} ms" for n, a, b in zip(names, t, t[1:])), flush=True)
