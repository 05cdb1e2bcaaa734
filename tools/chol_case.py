"""Direct-solver measurement: assemble + LinearSolver_Cholesky on a named workload, one JSON line.

    python tools/chol_case.py [workload] [reps]

Checks the solution by its residual through the product's SpMV (the CPU oracle would need hours at
this size) and against a CG solve to 1e-10.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "beam_100k_g2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if "x" in name:                                              # e.g. 40x40x60: an ad-hoc beam
    nx, ny, nz = (int(v) for v in name.split("x"))
    m = mesh.beam(nx, ny, nz, tolerance=1e-10)
else:
    m = mesh.workload(name, tolerance=1e-10)
with Solver() as s:
    s.SetModel(m)
    s.AssignDOF()
    a = s.ParallelAssembly_K()
    best = None
    for _ in range(reps):
        rep = s.LinearSolver_Cholesky()
        if best is None or rep.factor_ms < best.factor_ms:
            best = rep
    x = s.Exclude_BC_DOF()
    red = s.nDOF_reduction()
    U = np.zeros(m.n_dof)
    U[red >= 0] = x
    F = s.F()
    res = float(np.linalg.norm(s.spmv(U)[red >= 0] - F) / np.linalg.norm(F))
    cg = s.LinearSolver_CG(merit_check=0, IterMax=50000)
    xc = s.Exclude_BC_DOF()
    out = {
        "workload": name, "n_dof": int(best.n), "n_elem": m.n_elem, "block": best.block, "n_blocks": int(best.n_blocks),
        "skyline_gb": best.skyline_bytes / 1e9, "flops": best.flops, "setup_ms": best.setup_ms,
        "factor_ms": best.factor_ms, "solve_ms": best.solve_ms, "factor_tflops": best.flops / best.factor_ms / 1e9,
        "launches": int(best.kernel_launches), "terminationtype": best.terminationtype,
        "residual_rel": res, "cg_type": cg.terminationtype, "cg_its": cg.iterationscount, "cg_ms": cg.solve_ms,
        "vs_cg_rel": float(np.linalg.norm(xc - x) / np.linalg.norm(x)),
        "elements_per_s_solver_only": m.n_elem / ((best.setup_ms + best.factor_ms + best.solve_ms) / 1e3),
    }
    print(json.dumps(out))
