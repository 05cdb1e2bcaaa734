"""The native host end to end: STdb in -> stan_solver (GPU) -> STdb out, against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from stan_b200 import mesh, stdb

pytestmark = pytest.mark.gpu


def test_stan_solver_cli_roundtrip_against_oracle(oracle, tmp_path):
    from stan_b200 import build
    host = build.build_host()
    m = mesh.beam(5, 4, 24, jitter=True, n_parts=2, tolerance=1e-9)
    src, out = tmp_path / "model.STdb", tmp_path / "solved.STdb"
    src.write_bytes(stdb.encode(stdb.from_model(m)))
    r = subprocess.run([host, str(src), "-o", str(out), "--strict"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for line in ("Reading input file:", "DoF ordering:", "DATABASE SUMMARY", "LINEAR STATIC ANALYSIS", "K Matrix assembly:",
                 "Solving linear system...", "NORMAL  (type 1)", "Stress recovery:", "Total CPU time:"):
        assert line in r.stdout, line                            # console lines of Solver.cs / SolverFunctions.cs
    db = stdb.decode(out.read_bytes())
    ni, disp, strain, stress = stdb.results(db)
    o = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-9, merit_check=0, maxits=100000))
    assert np.array_equal(ni, o.node_index)                      # Node.DOF written back
    assert db.analysis.result_stepno == 1 and db.ndof == m.n_dof
    ou = o.U_full.reshape(-1, 3)[o.node_index]
    assert np.linalg.norm(disp - ou) / np.linalg.norm(ou) < 1e-9
    assert np.abs(stress - o.stress).max() <= 1e-7 * np.abs(o.stress).max()
    assert np.abs(strain - o.strain).max() <= 1e-7 * np.abs(o.strain).max()
    n0 = db.nodes[0]
    assert n0.dispx[0] == 0.0 and len(n0.dispx) == 2 and n0.elist == [1]          # [0, u]; EList of a corner node
    e0 = db.elems[0]
    assert len(e0.strain) == 2 and not any(e0.strain[0].M) and e0.strain[0].rows == 8 and e0.strain[0].cols == 6
    # solving the solved file again gives the same answer (results are overwritten, not appended)
    r2 = subprocess.run([host, str(out), "--strict"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    _, disp2, _, stress2 = stdb.results(stdb.decode(out.read_bytes()))
    assert np.array_equal(disp2, disp) and np.array_equal(stress2, stress)


def test_stan_solver_cli_cholesky(oracle, tmp_path):
    """Analysis.LinSolver = "Cholesky" takes the direct path (Solver.cs:163) and prints its console lines."""
    from stan_b200 import build
    host = build.build_host()
    m = mesh.beam(4, 4, 12, jitter=True)
    m.lin_solver = "Cholesky"
    src, out = tmp_path / "model.STdb", tmp_path / "solved.STdb"
    src.write_bytes(stdb.encode(stdb.from_model(m)))
    r = subprocess.run([host, str(src), "-o", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for line in ("Linear system K*U=F:", "- Cholesky decomposition:   Done", "NORMAL termination (type 1)",
                 "Total time to solve K*U=F:", "Stress recovery:"):
        assert line in r.stdout, line                            # SolverFunctions.cs:384-441
    db = stdb.decode(out.read_bytes())
    assert db.analysis.linsolver == "Cholesky"
    ni, disp, strain, stress = stdb.results(db)
    red, _ = oracle.spc_reduction(m, ni)
    K = oracle.assemble_upper(m, ni, red)
    xo, tt, _ = oracle.cholesky_skyline(K, oracle.build_rhs(m, ni, red))
    ou = oracle.include_bc_dof(red, xo).reshape(-1, 3)[ni]
    assert tt == 1 and np.abs(disp - ou).max() <= 1e-10 * np.abs(ou).max()
    # a solver name the path does not provide is refused, not silently replaced
    m.lin_solver = "LU"
    src.write_bytes(stdb.encode(stdb.from_model(m)))
    r = subprocess.run([host, str(src), "-o", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 3 and "LU" in r.stderr
