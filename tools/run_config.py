"""Runs one BASELINE.json configuration on one GPU and prints a JSON summary with
size-independent checks (true residual through stan_spmv, symmetry, fixed DOFs)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stan_b200 import mesh  # noqa: E402
from stan_b200.solver import Solver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "beam_1m_g1"
maxits = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
m = mesh.workload(name, tolerance=1e-8)
with Solver() as s:
    t0 = time.perf_counter()
    s.SetModel(m)
    ni = s.AssignDOF()
    s.ParallelAssembly_K()                # the first call pays pool growth and lazy module loading
    a = s.ParallelAssembly_K()
    cg =s.LinearSolver_CG(merit_check=0, IterMax=maxits)
    rc = s.Recovery_Stress()
    U = s.Include_BC_DOF()
    wall = time.perf_counter() - t0
    b = np.zeros(m.n_dof)
    np.add.at(b, 3 * ni[m.load_node], m.load_val[:, 0])
    r = b - s.spmv(U)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(m.n_dof), rng.standard_normal(m.n_dof)
    sym = abs(y @ s.spmv(x) - x @ s.spmv(y)) / abs(y @ s.spmv(x))
    strain, stress = s.strain_stress()
    print(json.dumps({
        "workload": name, "n_elem": m.n_elem, "n_dof": m.n_dof, "assembly_ms": a.total_ms, "ke_kernel_ms": a.assembly_ms,
        "assembly_el_s": m.n_elem / (a.total_ms * 1e-3), "cg_type": cg.terminationtype, "cg_iterations": cg.iterationscount,
        "cg_solve_ms": cg.solve_ms, "cg_iters_s": cg.iterationscount / (cg.solve_ms * 1e-3),
        "cg_rel_residual": float(np.sqrt(cg.r2) / cg.bnorm), "true_rel_residual": float(np.linalg.norm(r) / np.linalg.norm(b)),
        "symmetry_defect": float(sym), "fixed_dofs_zero": bool(not U.reshape(-1, 3)[ni[m.spc_node]].any()),
        "tip_ux": float(U[3 * ni[m.load_node]].mean()), "max_abs_stress": float(np.abs(stress).max()),
        "recovery_ms": rc.recover_ms, "wall_s": wall}))
