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
