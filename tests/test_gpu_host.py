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


def _vtu_arrays(path):
    """Minimal VTU reader for the two modes the writer produces (ascii, inline base64 with UInt32 headers)."""
    import base64
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    assert root.tag == "VTKFile" and root.get("type") == "UnstructuredGrid" and root.get("byte_order") == "LittleEndian"
    piece = root.find("UnstructuredGrid/Piece")
    dt = {"Float32": np.float32, "Int64": np.int64, "UInt8": np.uint8}
    out = {"n_points": int(piece.get("NumberOfPoints")), "n_cells": int(piece.get("NumberOfCells")), "order": []}
    for da in piece.iter("DataArray"):
        t, text = dt[da.get("type")], da.text.strip()
        if da.get("format") == "ascii":
            a = np.array(text.split(), dtype=np.float64).astype(t)
        else:
            n = int(np.frombuffer(base64.b64decode(text[:8]), dtype=np.uint32)[0])
            a = np.frombuffer(base64.b64decode(text[8:]), dtype=t)
            assert a.nbytes == n
        out[da.get("Name")] = a
        out["order"].append(da.get("Name"))
    return out


@pytest.mark.parametrize("ascii_mode", [False, True])
def test_stan_solver_cli_vtu_export(tmp_path, ascii_mode):
    """`--vtu` writes what PrePost's Export window writes (ExportWindow.xaml.cs:43-108, Part.ExportGrid
    Part.cs:857-939): per part its sorted node list as deformed points, hexahedra, nodal-averaged arrays."""
    from stan_b200 import build
    from stan_b200.solver import Solver
    host = build.build_host()
    m = mesh.beam(4, 3, 9, jitter=True, n_parts=3, tolerance=1e-10)
    src, out, prefix = tmp_path / "model.STdb", tmp_path / "solved.STdb", tmp_path / "res"
    src.write_bytes(stdb.encode(stdb.from_model(m)))
    cmd = [host, str(src), "-o", str(out), "--strict", "--vtu", str(prefix)] + (["--vtu-ascii"] if ascii_mode else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Result file exported" in r.stdout, r.stdout + r.stderr
    v = _vtu_arrays(str(prefix) + "_001.vtu")
    with Solver() as s:
        res = s.SolverLinearStatics(m, merit_check=0)
        _, point, _ = s.Load_Scalar()
    # expected layout: parts in order of appearance, each with its ascending distinct nodes
    pts, cells, srcs = [], [], []
    for pid in (1, 2, 3):
        el = np.where(m.elem_pid == pid)[0]
        ids = np.unique(m.conn[el])
        local = {int(n): len(srcs) + k for k, n in enumerate(ids)}
        srcs.extend(ids.tolist())
        cells.extend([[local[int(n)] for n in m.conn[e]] for e in el])
    srcs = np.array(srcs)
    assert v["n_points"] == len(srcs) > m.n_nodes and v["n_cells"] == m.n_elem      # interface nodes once per part
    assert np.array_equal(v["connectivity"].reshape(-1, 8), np.array(cells))
    assert np.array_equal(v["offsets"], 8 * np.arange(1, m.n_elem + 1)) and np.all(v["types"] == 12)
    np.testing.assert_allclose(v["Points"].reshape(-1, 3), (m.xyz + res.disp)[srcs].astype(np.float32), rtol=2e-6, atol=1e-7)
    names = ["Displacement X", "Displacement Y", "Displacement Z", "Total Displacement", "Stress XX", "Stress YY", "Stress ZZ",
             "Stress XY", "Stress YZ", "Stress XZ", "Stress P1", "Stress P2", "Stress P3", "von Mises Stress", "Strain XX",
             "Strain YY", "Strain ZZ", "Strain XY", "Strain YZ", "Strain XZ", "Strain P1", "Strain P2", "Strain P3",
             "Effective Strain"]
    assert v["order"][:24] == names[:4] + names[14:] + names[4:14]                   # Displacement, Strain, Stress
    for k, name in enumerate(names):
        scale = np.abs(point[:, k]).max() + 1e-30
        assert np.abs(v[name] - point[srcs, k]).max() <= 2e-6 * scale, name
