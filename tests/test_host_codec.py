"""Native host pieces that need no GPU: STdb codec (two independent implementations against each
other and against hand-assembled wire bytes) and the Nastran import (Database.ReadNastranMesh)."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from stan_b200 import mesh, stdb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    from stan_b200 import build
    return build.build_host()


def _run(host, *args):
    r = subprocess.run([host, *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_wire_known_answer():
    # Database{NodeLib: {7: Node{ID=7, X=1.5, DOF=[0,1,2]}}, nDOF=3} assembled by hand from the protobuf spec
    node = b"\x08\x07" + b"\x11" + struct.pack("<d", 1.5) + b"\x30\x00\x30\x01\x30\x02"
    entry = b"\x08\x07" + b"\x12" + bytes([len(node)]) + node
    golden = b"\x0a" + bytes([len(entry)]) + entry + b"\x28\x03"
    db = stdb.Database(nodes=[stdb.Node(id=7, x=1.5, dof=[0, 1, 2])], ndof=3)
    assert stdb.encode(db) == golden
    back = stdb.decode(golden)
    assert back.nodes[0].id == 7 and back.nodes[0].x == 1.5 and back.nodes[0].y == 0.0 and back.nodes[0].dof == [0, 1, 2]
    assert back.ndof == 3
    # negative int32 is a 10-byte varint; packed and unpacked repeated scalars decode alike
    neg = stdb.decode(stdb.encode(stdb.Database(nodes=[stdb.Node(id=-5, elist=[-1, 3])])))
    assert neg.nodes[0].id == -5 and neg.nodes[0].elist == [-1, 3]
    packed_node = b"\x08\x07" + b"\x32\x03\x00\x01\x02" + b"\x3a\x10" + struct.pack("<dd", 0.0, 2.5)
    e2 = b"\x08\x07\x12" + bytes([len(packed_node)]) + packed_node
    p = stdb.decode(b"\x0a" + bytes([len(e2)]) + e2)
    assert p.nodes[0].dof == [0, 1, 2] and p.nodes[0].dispx == [0.0, 2.5]


def test_python_and_native_codec_agree(host, tmp_path):
    m = mesh.beam(3, 2, 4, jitter=True, n_parts=2, tolerance=1e-7, max_iter=123)
    db = stdb.from_model(m)
    db.info_raw = b"\x0a\x06\x08\x01\x12\x02\x08\x03"            # opaque Information payload must survive
    db.elems[0].strain = [stdb.MatrixST([0.0] * 48, 8, 6), stdb.MatrixST(list(np.arange(48.0) - 7), 8, 6)]
    raw = stdb.encode(db)
    a, b = tmp_path / "a.STdb", tmp_path / "b.STdb"
    a.write_bytes(raw)
    _run(host, "--roundtrip", str(a), str(b))
    assert b.read_bytes() == raw                                 # byte-identical re-serialisation
    d = json.loads(_run(host, "--dump", str(a)))
    assert d["nodes"] == m.n_nodes and d["elements"] == m.n_elem and d["materials"] == 2 and d["bcs"] == 2
    assert d["ndof"] == m.n_dof and d["analysis"] == "Linear_Statics" and d["linsolver"] == "CG"
    assert d["tolerance"] == 1e-7 and d["itermax"] == 123 and d["result_stepno"] == 0
    back = stdb.decode(raw)
    assert [n.id for n in back.nodes] == list(range(1, m.n_nodes + 1))
    assert back.elems[5].nlist == [int(v) + 1 for v in m.conn[5]] and back.elems[5].matid == int(m.elem_mat[5]) + 1
    assert back.bcs[0][1].type == "SPC" and back.bcs[1][1].nodal[0][1].M == list(m.load_val[0])
    assert back.info_raw == db.info_raw and back.elems[0].strain[1].array()[7, 5] == 40.0


def test_truncated_file_is_an_error(host, tmp_path):
    raw = stdb.encode(stdb.from_model(mesh.beam(1, 1, 1)))
    p = tmp_path / "t.STdb"
    p.write_bytes(raw[: len(raw) // 2])
    r = subprocess.run([host, "--dump", str(p)], capture_output=True, text=True)
    assert r.returncode != 0 and "malformed" in r.stderr.lower() or "truncated" in r.stderr.lower()


def test_bdf_import_readme_excerpt(host, tmp_path):
    out = tmp_path / "m.STdb"
    rep = json.loads(_run(host, "--import-bdf", os.path.join(ROOT, "tests", "golden", "readme_excerpt.bdf"), str(out)))
    assert rep == {"nodes": 3, "elements": 3, "import_errors": 0}
    db = stdb.decode(out.read_bytes())
    assert [(n.id, n.x, n.y, n.z) for n in db.nodes] == [(1, 0.0, 15.0, 0.0), (2, -7.11e-15, 5.0, 0.0), (3, 0.0, -5.0, 0.0)]
    assert db.elems[0].nlist == [573, 570, 571, 572, 1236, 1237, 1238, 1239] and db.elems[0].pid == 1
    assert db.elems[2].nlist == [576, 573, 572, 574, 1242, 1236, 1239, 1243]
    assert all(e.type == "HEX8_G2" and e.matid == 0 for e in db.elems)       # Element.cs:59, :64
    assert db.ndof == 9 and db.analysis.linsolver == "CG" and db.analysis.tolerance == 1e-6


def test_bdf_import_of_generated_mesh(host, tmp_path):
    m = mesh.beam(3, 2, 2)
    bdf, out = tmp_path / "g.bdf", tmp_path / "g.STdb"
    mesh.write_bdf(m, str(bdf))
    rep = json.loads(_run(host, "--import-bdf", str(bdf), str(out)))
    assert rep["nodes"] == m.n_nodes and rep["elements"] == m.n_elem and rep["import_errors"] == 0
    db = stdb.decode(out.read_bytes())
    np.testing.assert_array_equal(np.array([[n.x, n.y, n.z] for n in db.nodes]), m.xyz)
    np.testing.assert_array_equal(np.array([e.nlist for e in db.elems]), m.conn + 1)
    # quirks of the reference parser: '+' exponents are not patched (node dropped), duplicate ids are dropped
    bad = tmp_path / "bad.bdf"
    bad.write_text("GRID           1             0.0     1.0     2.0\nGRID           1             9.0     9.0     9.0\n"
                   "GRID           2         7.11+15     5.0     0.0\nCTETRA         9       1       1       2       3       4\n")
    rep = json.loads(_run(host, "--import-bdf", str(bad), str(out)))
    assert rep == {"nodes": 1, "elements": 0, "import_errors": 2}


def test_build_model_from_bdf_and_pasted_bc_text(host, tmp_path):
    """`--build`: the PrePost steps between import and solve (README.md:50-76 of the reference) scripted —
    materials (MainWindow.AddMat), part material / element type (BOX_Part), boundary conditions in the
    Paste format with its quirks (BOX_BC.xaml.cs:228-270, BoundaryCondition.cs:87-98), Analysis settings."""
    m = mesh.beam(3, 2, 5, n_parts=2)
    bdf, out = tmp_path / "mesh.bdf", tmp_path / "model.STdb"
    mesh.write_bdf(m, str(bdf))
    spc = tmp_path / "spc.txt"
    spc.write_text("1\t1\t1\t1\r\n2\t1\t0\t1\r\nnot a row\n3,1,1,0\n4 0 0 1\n9999\t1\t1\t1\n5\t1\t1\n6\t1e0\t1.0\t+1\n")
    load = tmp_path / "load.txt"
    load.write_text("70\t12.5\t0\t-3e2\n71\t0.5\t0\t0\n")
    r = subprocess.run([host, "--build", str(bdf), str(out), "--material", "210000", "0.3", "--material", "70000", "0.33",
                        "--part-mat", "1", "1", "--part-mat", "2", "2", "--elem-type", "HEX8_G1", "--spc", str(spc),
                        "--load", str(load), "--solver", "Cholesky", "--tol", "1e-9", "--itermax", "500"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout)
    assert info == {"nodes": m.n_nodes, "elements": m.n_elem, "import_errors": 0, "materials": 2, "bcs": 2, "bc_rows": 7}
    db = stdb.decode(out.read_bytes())
    assert [(x.id, x.type, x.name, x.E, x.poisson, x.colorid) for x in db.mats] == \
        [(1, "Elastic", "New Material", 210000.0, 0.3, 1), (2, "Elastic", "New Material", 70000.0, 0.33, 2)]
    assert all(e.type == "HEX8_G1" for e in db.elems)
    assert [e.matid for e in db.elems] == [int(p) for p in m.elem_pid]             # part 1 -> material 1, part 2 -> 2
    (k1, fix), (k2, ld) = db.bcs
    assert (k1, fix.id, fix.type, fix.colorid, k2, ld.id, ld.type, ld.colorid) == (1, 1, "SPC", 1, 2, 2, "PointLoad", 2)
    # tab, comma and space rows are read; CRLF tolerated; malformed / 3-field / unknown-node rows are dropped
    assert [(n, v.M, v.rows, v.cols) for n, v in fix.nodal] == \
        [(1, [1.0, 1.0, 1.0], 3, 1), (2, [1.0, 0.0, 1.0], 3, 1), (3, [1.0, 1.0, 0.0], 3, 1), (4, [0.0, 0.0, 1.0], 3, 1),
         (6, [1.0, 1.0, 1.0], 3, 1)]
    assert [(n, v.M) for n, v in ld.nodal] == [(70, [12.5, 0.0, -300.0]), (71, [0.5, 0.0, 0.0])]
    a = db.analysis
    assert (a.type, a.linsolver, a.tolerance, a.itermax, a.result_stepno) == ("Linear_Statics", "Cholesky", 1e-9, 500, 0)
    assert db.ndof == m.n_dof
    # a single pasted line is ignored like the reference's `text.Length > 1`; a repeated node is its Dictionary.Add failure
    one = tmp_path / "one.txt"
    one.write_text("1\t1\t1\t1")
    r = subprocess.run([host, "--build", str(bdf), str(out), "--spc", str(one)], capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout)["bc_rows"] == 0
    assert stdb.decode(out.read_bytes()).analysis.linsolver == "CG" and stdb.decode(out.read_bytes()).analysis.tolerance == 1e-6
    dup = tmp_path / "dup.txt"
    dup.write_text("1\t1\t1\t1\n1\t0\t0\t0\n")
    r = subprocess.run([host, "--build", str(bdf), str(out), "--spc", str(dup)], capture_output=True, text=True)
    assert r.returncode == 3 and "listed twice" in r.stderr


def test_remove_results(host, tmp_path):
    """MainWindow.RemoveResults_Click (MainWindow.xaml.cs:731-763): Result_StepNo = 0, Element.ClearResults,
    Node.Initialize_StepZero; numbering and everything else stays."""
    m = mesh.beam(2, 2, 3)
    db = stdb.from_model(m)
    for k, n in enumerate(db.nodes):
        n.dof = [3 * k, 3 * k + 1, 3 * k + 2]; n.elist = [1]
        n.dispx, n.dispy, n.dispz = [0.0, 0.5 * k], [0.0, -1.0], [0.0, 2.0]
    z = stdb.MatrixST([0.0] * 48, 8, 6)
    for e in db.elems:
        e.strain = [z, stdb.MatrixST([1e-3] * 48, 8, 6)]; e.stress = [z, stdb.MatrixST([210.0] * 48, 8, 6)]
    db.analysis.result_stepno = 1
    src, out = tmp_path / "solved.STdb", tmp_path / "clean.STdb"
    src.write_bytes(stdb.encode(db))
    r = subprocess.run([host, "--remove-results", str(src), str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert out.stat().st_size < src.stat().st_size / 2
    c = stdb.decode(out.read_bytes())
    assert c.analysis.result_stepno == 0 and c.analysis.tolerance == db.analysis.tolerance
    assert all(not e.strain and not e.stress for e in c.elems)
    assert all(n.dispx == [0.0] and n.dispy == [0.0] and n.dispz == [0.0] for n in c.nodes)
    assert [n.dof for n in c.nodes] == [n.dof for n in db.nodes] and [e.nlist for e in c.elems] == [e.nlist for e in db.elems]
    assert [[(n, v.M) for n, v in bc.nodal] for _, bc in c.bcs] == [[(n, v.M) for n, v in bc.nodal] for _, bc in db.bcs]


def test_codec_fuzz_python_vs_native(host, tmp_path):
    """Random databases (hypothesis): the Python codec is its own inverse and the independent C++ codec
    re-serialises every one of them byte for byte — negative ids, empty and long lists, odd doubles,
    non-ASCII names, results present or absent."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    i32 = st.integers(-2**31, 2**31 - 1)
    f64 = st.floats(allow_nan=False, width=64)
    text = st.text(max_size=12)
    mat_st = st.builds(lambda r, c, fill: stdb.MatrixST([fill * (k + 1) for k in range(r * c)], r, c),
                       st.integers(0, 4), st.integers(0, 4), f64)
    node = st.builds(stdb.Node, id=i32, x=f64, y=f64, z=f64, elist=st.lists(i32, max_size=5), dof=st.lists(i32, max_size=3),
                     dispx=st.lists(f64, max_size=3), dispy=st.lists(f64, max_size=3), dispz=st.lists(f64, max_size=3))
    elem = st.builds(stdb.Element, id=i32, type=text, pid=i32, matid=i32, nlist=st.lists(i32, max_size=8),
                     strain=st.lists(mat_st, max_size=2), stress=st.lists(mat_st, max_size=2))
    mat = st.builds(stdb.Material, id=i32, type=text, name=text, E=f64, poisson=f64, colorid=i32)
    bc = st.tuples(i32, st.builds(stdb.BoundaryCondition, type=text, name=text, id=i32, colorid=i32,
                                  nodal=st.lists(st.tuples(i32, mat_st), max_size=4)))
    ana = st.one_of(st.none(), st.builds(stdb.Analysis, type=text, linsolver=text, tolerance=f64, itermax=i32, incnumb=i32,
                                         result_stepno=i32))
    db_st = st.builds(stdb.Database, nodes=st.lists(node, max_size=4), elems=st.lists(elem, max_size=3),
                      mats=st.lists(mat, max_size=2), bcs=st.lists(bc, max_size=2), ndof=i32, analysis=ana,
                      info_raw=st.one_of(st.none(), st.binary(max_size=6)))
    a, b = tmp_path / "f.STdb", tmp_path / "g.STdb"

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(db_st)
    def run(db):
        raw = stdb.encode(db)
        assert stdb.encode(stdb.decode(raw)) == raw
        a.write_bytes(raw)
        r = subprocess.run([host, "--roundtrip", str(a), str(b)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        assert b.read_bytes() == raw

    run()


def test_bdf_import_fuzz_against_second_restatement(host, tmp_path):
    """Generated decks with the oddities real decks have (exponent-less floats, blank fields, continuation
    lines, '+' markers, comments, CRLF, duplicates, junk): the native importer and the independent Python
    restatement of the reference's parser (oracle/bdf_import.py) must agree on every node, element and
    on the number of import errors."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    from oracle import bdf_import

    num = st.sampled_from(["1.5", "-2.5", ".5", "-.25", "1.-3", "-1.5-3", "7.11-15", "7.11+15", "1e-3", "2.E+2", "0.0", "12.",
                           "", "abc", "1.2.3", "-", "1-", "3", "-7", "1e", "+4.5", "1,5"])
    ident = st.one_of(st.integers(1, 40).map(str), st.sampled_from(["", "x1", "-3", "99999999", "123456789", "1.0", "+7"]))
    field = lambda s: s.map(lambda v: v[:8].rjust(8))                       # noqa: E731
    grid = st.tuples(field(ident), field(st.sampled_from(["", "", "", "0"])), field(num), field(num), field(num),
                     st.sampled_from(["", "       0", "  ", "\t"])).map(lambda t: "GRID    " + "".join(t))
    nid = st.one_of(st.integers(1, 60).map(str), st.sampled_from(["", "q", "+", "1+", "2147483648"]))
    chexa = st.tuples(st.sampled_from(["CHEXA   ", "CHEXA  ", " CHEXA   ", "$CHEXA  ", "XCHEXA  "]), field(ident), field(ident),
                      st.lists(field(nid), min_size=0, max_size=6), st.sampled_from(["+", "+E1", "", " "]),
                      st.sampled_from(["+       ", "+E1     ", "        ", "*       ", ""]), st.lists(field(nid), min_size=0, max_size=3)) \
        .map(lambda t: t[0] + t[1] + t[2] + "".join(t[3]) + t[4] + "\n" + t[5] + "".join(t[6]))
    other = st.sampled_from(["$$ comment", "", "ENDDATA", "CTETRA         9       1       1       2       3       4", "GRIDX  1",
                             "BEGIN BULK", " GRID          1             0.0     0.0     0.0", "+ stray continuation"])
    deck = st.tuples(st.lists(st.one_of(grid, chexa, other), min_size=0, max_size=14), st.sampled_from(["\n", "\r\n"]),
                     st.booleans()).map(lambda t: t[1].join(l.replace("\n", t[1]) for l in t[0]) + (t[1] if t[2] else ""))
    bdf, out = tmp_path / "f.bdf", tmp_path / "f.STdb"

    @settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(deck)
    def run(text):
        bdf.write_bytes(text.encode())
        r = subprocess.run([host, "--import-bdf", str(bdf), str(out)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        rep = json.loads(r.stdout)
        nodes, elems, errors = bdf_import.read_nastran_mesh(text)
        db = stdb.decode(out.read_bytes())
        assert [(n.id, n.x, n.y, n.z) for n in db.nodes] == [(k, *v) for k, v in nodes.items()]
        assert [(e.id, e.pid, e.nlist, e.type) for e in db.elems] == [(k, v[0], v[1], v[2]) for k, v in elems.items()]
        assert rep["import_errors"] == errors and rep["nodes"] == len(nodes) and rep["elements"] == len(elems)

    run()


def test_bc_paste_fuzz_against_second_restatement(host, tmp_path):
    """Pasted boundary-condition text: native `--build` against oracle/bdf_import.parse_bc_text / apply_bc."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    from oracle import bdf_import

    m = mesh.beam(1, 1, 2)                                                   # nodes 1..12
    bdf, out, txt = tmp_path / "m.bdf", tmp_path / "m.STdb", tmp_path / "bc.txt"
    mesh.write_bdf(m, str(bdf))
    tok = st.sampled_from(["1", "2", "3", "7", "12", "13", "0", "-1", "1.0", "1e0", "0.5", "-2.5e-3", "+4", ".5", "5.", "", "x",
                           "1 ", " 1", "1\r", "99999999999", "1e", "--1"])
    sep = st.sampled_from([",", " ", "\t", "\t", "\t"])
    row = st.tuples(st.lists(tok, min_size=2, max_size=5), sep).map(lambda t: t[1].join(t[0]))
    text = st.tuples(st.lists(row, min_size=0, max_size=8), st.sampled_from(["\n", "\r\n"]), st.booleans()) \
        .map(lambda t: t[1].join(t[0]) + (t[1] if t[2] else ""))

    @settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
    @given(text)
    def run(s):
        txt.write_bytes(s.encode())
        r = subprocess.run([host, "--build", str(bdf), str(out), "--spc", str(txt)], capture_output=True, text=True, timeout=60)
        try:
            want = bdf_import.apply_bc(bdf_import.parse_bc_text(s), set(range(1, m.n_nodes + 1)))
        except KeyError:
            assert r.returncode == 3 and "listed twice" in r.stderr
            return
        assert r.returncode == 0, r.stderr
        (_, bc), = stdb.decode(out.read_bytes()).bcs
        assert [(n, v.M) for n, v in bc.nodal] == want

    run()


def test_to_model_inverts_from_model():
    m = mesh.beam(3, 2, 4, jitter=True, n_parts=2, tolerance=3e-7, max_iter=77, elem_type=mesh.HEX8_G1)
    m.lin_solver = "Cholesky"
    back = stdb.to_model(stdb.decode(stdb.encode(stdb.from_model(m))))
    for f in ("xyz", "conn", "elem_type", "elem_mat", "elem_pid", "mat_E", "mat_nu", "spc_node", "spc_val", "load_node", "load_val"):
        a, b = getattr(m, f), getattr(back, f)
        assert a.dtype == b.dtype and np.array_equal(a, b), f
    assert (back.tolerance, back.max_iter, back.lin_solver) == (3e-7, 77, "Cholesky")
    db = stdb.from_model(m)
    db.elems[0].type = "TET4_G2"
    with pytest.raises(ValueError):
        stdb.to_model(db)
