"""How far does summation order alone move a Jacobi-CG trajectory?  ("1007 vs 1067 on the 100k beam")

Runs the CPU restatement of alglib.lincg (oracle/stan_oracle.c, SolverFunctions.cs:270-330) on one matrix in four
equally valid roundings of the same recurrences — symmetric product in ALGLIB's sparsesmv order or row-wise over
the expanded matrix, dot products left-to-right or in 1024-term blocks — and, when a GPU is present, libstan_b200
on the same matrix.  Prints iteration counts, the first iteration at which ||r_k||^2 / alpha_k / beta_k differ
from the ALGLIB-order run by more than 1e-9 / 1e-6 / 1e-3, and the distance between the converged solutions.

    python tools/cg_trajectory.py [workload|nx,ny,nz[,jitter]] [epsf] > profiles/r02_cg_trajectory_<case>.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from stan_b200 import mesh  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "beam_100k_g2"
epsf = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
if case in mesh.WORKLOADS:
    m = mesh.workload(case, tolerance=epsf)
else:
    f = case.split(",")
    m = mesh.beam(int(f[0]), int(f[1]), int(f[2]), jitter=len(f) > 3 and f[3] == "jitter", tolerance=epsf)
O.set_threads()
ni = O.assign_dof(m)
red, _ = O.spc_reduction(m, ni)
F = O.build_rhs(m, ni, red)
K = O.assemble_upper(m, ni, red)
CAP = 20000


def first_above(rel, tol):
    idx = np.nonzero(rel > tol)[0]
    return int(idx[0]) + 1 if idx.size else None


runs = {}
for name, ps, dm in (("alglib_order", 0, 0), ("alglib_spmv_blocked_dots", 0, 1), ("rowwise_spmv", 1, 0),
                     ("rowwise_spmv_blocked_dots", 1, 1)):
    t0 = time.perf_counter()
    x, rep, h = O.lincg_history(K, F, O.cg_opts(epsf=epsf, merit_check=0, maxits=CAP, parallel_spmv=ps, dot_mode=dm), CAP)
    runs[name] = (x, rep.iterationscount, rep.terminationtype, h, time.perf_counter() - t0)

try:
    import torch
    have_gpu = torch.cuda.is_available()
except Exception:
    have_gpu = False
if have_gpu:
    from stan_b200.solver import Solver
    with Solver() as s:
        s.SetModel(m); s.SetDOF(ni); s.ParallelAssembly_K()
        s.cg_history(CAP)
        rep = s.LinearSolver_CG(tolerance=epsf, merit_check=0, IterMax=CAP)
        runs["libstan_b200"] = (s.Exclude_BC_DOF(), rep.iterationscount, rep.terminationtype, s.cg_history(), rep.solve_ms * 1e-3)

x0, _, _, h0, _ = runs["alglib_order"]
out = {"case": case, "epsf": epsf, "n_free": int(K.n), "nnz_upper": int(K.nnz), "runs": {}}
for name, (x, its, tt, h, sec) in runs.items():
    n = min(len(h), len(h0))
    rel = (np.abs(h[:n, :3] - h0[:n, :3]) / np.maximum(np.abs(h0[:n, :3]), 1e-300)).max(axis=1)
    out["runs"][name] = {"iterations": int(its), "terminationtype": int(tt), "seconds": round(sec, 2),
                         "first_iteration_off_by": {"1e-9": first_above(rel, 1e-9), "1e-6": first_above(rel, 1e-6),
                                                    "1e-3": first_above(rel, 1e-3)},
                         "max_rel_diff_first_40": float(rel[:40].max()) if n >= 40 else None,
                         "solution_rel_diff_vs_alglib_order": float(np.linalg.norm(x - x0) / np.linalg.norm(x0))}
print(json.dumps(out, indent=1))
