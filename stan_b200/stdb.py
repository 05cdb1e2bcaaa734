"""STdb (STAN database) in Python: protobuf wire format of the [ProtoContract] classes in
/root/reference/src/STAN_Database.

An independent second implementation of the codec in stan_b200/host/stdb.cpp: tests write a
database from a flat `Model`, let the native host (stan_b200/lib/stan_solver) solve it, and read
the results back with this module.  No protobuf runtime is needed: seven small messages,
Dictionary<int,T> = repeated {1: key, 2: value}, int = varint, double = fixed64, zero/empty members
omitted, repeated scalars accepted packed or unpacked and written unpacked.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from .mesh import HEX8_G1, HEX8_G2, Model

# ---------------------------------------------------------------------------- wire primitives


def _varint(v: int) -> bytes:
    v &= 0xFFFFFFFFFFFFFFFF                      # negative int32 -> 10-byte two's complement
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _tag(f: int, wt: int) -> bytes:
    return _varint((f << 3) | wt)


def _i32(f: int, v: int, force=False) -> bytes:
    return b"" if (v == 0 and not force) else _tag(f, 0) + _varint(int(v))


def _f64(f: int, v: float, force=False) -> bytes:
    return b"" if (v == 0.0 and not force) else _tag(f, 1) + struct.pack("<d", float(v))


def _str(f: int, s: str) -> bytes:
    b = s.encode("utf-8")
    return b"" if not b else _tag(f, 2) + _varint(len(b)) + b


def _msg(f: int, payload: bytes) -> bytes:
    return _tag(f, 2) + _varint(len(payload)) + payload


def _rep_f64(f: int, values) -> bytes:
    a = np.asarray(values, dtype="<f8").ravel()
    if a.size == 0:
        return b""
    rec = np.zeros(a.size, dtype=[("t", "u1"), ("v", "<f8")])   # tag (f <= 15 fits one byte) + fixed64, unpacked
    rec["t"] = (f << 3) | 1
    rec["v"] = a
    return rec.tobytes()


def _rep_i32(f: int, values) -> bytes:
    return b"".join(_tag(f, 0) + _varint(int(v)) for v in values)


class _Reader:
    def __init__(self, buf: bytes, pos: int = 0, end: int | None = None):
        self.b, self.p, self.e = buf, pos, len(buf) if end is None else end

    def more(self):
        return self.p < self.e

    def varint(self) -> int:
        v = shift = 0
        while True:
            c = self.b[self.p]
            self.p += 1
            v |= (c & 0x7F) << shift
            if not c & 0x80:
                return v
            shift += 7

    def sint32(self) -> int:
        v = self.varint() & 0xFFFFFFFF
        return v - (1 << 32) if v & 0x80000000 else v

    def f64(self) -> float:
        v = struct.unpack_from("<d", self.b, self.p)[0]
        self.p += 8
        return v

    def sub(self) -> "_Reader":
        n = self.varint()
        r = _Reader(self.b, self.p, self.p + n)
        self.p += n
        return r

    def raw(self) -> bytes:
        n = self.varint()
        out = self.b[self.p:self.p + n]
        self.p += n
        return out

    def tag(self):
        t = self.varint()
        return t >> 3, t & 7

    def skip(self, wt):
        if wt == 0:
            self.varint()
        elif wt == 1:
            self.p += 8
        elif wt == 2:
            self.raw()
        elif wt == 5:
            self.p += 4
        else:
            raise ValueError(f"bad wire type {wt}")

    def rep_i32(self, wt, out):
        if wt == 2:
            r = self.sub()
            while r.more():
                out.append(r.sint32())
        else:
            out.append(self.sint32())

    def rep_f64(self, wt, out):
        if wt == 2:
            raw = self.raw()
            out.extend(np.frombuffer(raw, dtype="<f8").tolist())
        else:
            out.append(self.f64())


# ---------------------------------------------------------------------------- object model


@dataclass
class MatrixST:
    M: list = field(default_factory=list)
    rows: int = 0
    cols: int = 0

    def encode(self) -> bytes:
        return _rep_f64(1, self.M) + _i32(2, self.rows) + _i32(3, self.cols)

    @staticmethod
    def decode(r: _Reader) -> "MatrixST":
        m = MatrixST()
        while r.more():
            f, wt = r.tag()
            if f == 1:
                r.rep_f64(wt, m.M)
            elif f == 2:
                m.rows = r.sint32()
            elif f == 3:
                m.cols = r.sint32()
            else:
                r.skip(wt)
        return m

    def array(self) -> np.ndarray:
        return np.asarray(self.M, dtype=float).reshape(self.rows, self.cols)


@dataclass
class Node:
    id: int = 0
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0
    elist: list = field(default_factory=list)
    dof: list = field(default_factory=list)
    dispx: list = field(default_factory=list)
    dispy: list = field(default_factory=list)
    dispz: list = field(default_factory=list)


@dataclass
class Element:
    id: int = 0
    type: str = ""
    pid: int = 0
    matid: int = 0
    nlist: list = field(default_factory=list)
    strain: list = field(default_factory=list)
    stress: list = field(default_factory=list)


@dataclass
class Material:
    id: int = 0
    type: str = ""
    name: str = ""
    E: float = 0.0
    poisson: float = 0.0
    colorid: int = 0


@dataclass
class BoundaryCondition:
    type: str = ""
    name: str = ""
    id: int = 0
    nodal: list = field(default_factory=list)     # [(node id, MatrixST 3x1)]
    colorid: int = 0


@dataclass
class Analysis:
    type: str = "Linear_Statics"
    linsolver: str = "CG"
    tolerance: float = 1.0e-6
    itermax: int = 0
    incnumb: int = 0
    result_stepno: int = 0


@dataclass
class Database:
    nodes: list = field(default_factory=list)
    elems: list = field(default_factory=list)
    mats: list = field(default_factory=list)
    bcs: list = field(default_factory=list)       # [(key, BoundaryCondition)]
    ndof: int = 0
    analysis: Analysis | None = None
    info_raw: bytes | None = None


def _entry(key: int, payload: bytes) -> bytes:
    return _i32(1, key) + _msg(2, payload)


def encode(db: Database) -> bytes:
    out = []
    for n in db.nodes:
        p = (_i32(1, n.id) + _f64(2, n.x) + _f64(3, n.y) + _f64(4, n.z) + _rep_i32(5, n.elist) + _rep_i32(6, n.dof)
             + _rep_f64(7, n.dispx) + _rep_f64(8, n.dispy) + _rep_f64(9, n.dispz))
        out.append(_msg(1, _entry(n.id, p)))
    for e in db.elems:
        p = (_i32(1, e.id) + _str(2, e.type) + _i32(3, e.pid) + _i32(4, e.matid) + _rep_i32(5, e.nlist)
             + b"".join(_msg(6, m.encode()) for m in e.strain) + b"".join(_msg(7, m.encode()) for m in e.stress))
        out.append(_msg(2, _entry(e.id, p)))
    for m in db.mats:
        p = _i32(1, m.id) + _str(2, m.type) + _str(3, m.name) + _f64(4, m.E) + _f64(5, m.poisson) + _i32(6, m.colorid)
        out.append(_msg(3, _entry(m.id, p)))
    for key, b in db.bcs:
        p = (_str(1, b.type) + _str(2, b.name) + _i32(3, b.id)
             + b"".join(_msg(4, _entry(nid, mat.encode())) for nid, mat in b.nodal) + _i32(5, b.colorid))
        out.append(_msg(4, _entry(key, p)))
    out.append(_i32(5, db.ndof))
    if db.analysis is not None:
        a = db.analysis
        out.append(_msg(6, _str(1, a.type) + _str(2, a.linsolver) + _f64(3, a.tolerance) + _i32(4, a.itermax)
                        + _i32(5, a.incnumb) + _i32(6, a.result_stepno)))
    if db.info_raw is not None:
        out.append(_msg(7, db.info_raw))
    return b"".join(out)


def _read_entry(r: _Reader):
    key, val = 0, None
    while r.more():
        f, wt = r.tag()
        if f == 1:
            key = r.sint32()
        elif f == 2:
            val = r.sub()
        else:
            r.skip(wt)
    return key, val


def decode(buf: bytes) -> Database:
    db = Database()
    r = _Reader(buf)
    while r.more():
        f, wt = r.tag()
        if f == 1:
            key, v = _read_entry(r.sub())
            n = Node(id=key)
            while v is not None and v.more():
                g, w = v.tag()
                if g == 1: n.id = v.sint32()
                elif g == 2: n.x = v.f64()
                elif g == 3: n.y = v.f64()
                elif g == 4: n.z = v.f64()
                elif g == 5: v.rep_i32(w, n.elist)
                elif g == 6: v.rep_i32(w, n.dof)
                elif g == 7: v.rep_f64(w, n.dispx)
                elif g == 8: v.rep_f64(w, n.dispy)
                elif g == 9: v.rep_f64(w, n.dispz)
                else: v.skip(w)
            db.nodes.append(n)
        elif f == 2:
            key, v = _read_entry(r.sub())
            e = Element(id=key)
            while v is not None and v.more():
                g, w = v.tag()
                if g == 1: e.id = v.sint32()
                elif g == 2: e.type = v.raw().decode("utf-8")
                elif g == 3: e.pid = v.sint32()
                elif g == 4: e.matid = v.sint32()
                elif g == 5: v.rep_i32(w, e.nlist)
                elif g == 6: e.strain.append(MatrixST.decode(v.sub()))
                elif g == 7: e.stress.append(MatrixST.decode(v.sub()))
                else: v.skip(w)
            db.elems.append(e)
        elif f == 3:
            key, v = _read_entry(r.sub())
            m = Material(id=key)
            while v is not None and v.more():
                g, w = v.tag()
                if g == 1: m.id = v.sint32()
                elif g == 2: m.type = v.raw().decode("utf-8")
                elif g == 3: m.name = v.raw().decode("utf-8")
                elif g == 4: m.E = v.f64()
                elif g == 5: m.poisson = v.f64()
                elif g == 6: m.colorid = v.sint32()
                else: v.skip(w)
            db.mats.append(m)
        elif f == 4:
            key, v = _read_entry(r.sub())
            b = BoundaryCondition()
            while v is not None and v.more():
                g, w = v.tag()
                if g == 1: b.type = v.raw().decode("utf-8")
                elif g == 2: b.name = v.raw().decode("utf-8")
                elif g == 3: b.id = v.sint32()
                elif g == 4:
                    nid, mv = _read_entry(v.sub())
                    b.nodal.append((nid, MatrixST.decode(mv) if mv is not None else MatrixST()))
                elif g == 5: b.colorid = v.sint32()
                else: v.skip(w)
            db.bcs.append((key, b))
        elif f == 5:
            db.ndof = r.sint32()
        elif f == 6:
            v = r.sub()
            a = Analysis(type="", linsolver="", tolerance=0.0)
            while v.more():
                g, w = v.tag()
                if g == 1: a.type = v.raw().decode("utf-8")
                elif g == 2: a.linsolver = v.raw().decode("utf-8")
                elif g == 3: a.tolerance = v.f64()
                elif g == 4: a.itermax = v.sint32()
                elif g == 5: a.incnumb = v.sint32()
                elif g == 6: a.result_stepno = v.sint32()
                else: v.skip(w)
            db.analysis = a
        elif f == 7:
            db.info_raw = r.raw()
        else:
            r.skip(wt)
    return db


# ---------------------------------------------------------------------------- Model <-> Database


def from_model(m: Model, *, first_id: int = 1) -> Database:
    """What PrePost would save for this model: 1-based IDs, one SPC and one PointLoad BC,
    materials 1..n, Analysis = Linear_Statics / CG / tolerance / max_iter."""
    db = Database()
    for i, (x, y, z) in enumerate(m.xyz):
        db.nodes.append(Node(id=first_id + i, x=float(x), y=float(y), z=float(z), dof=[0, 0, 0], dispx=[0.0], dispy=[0.0], dispz=[0.0]))
    for e in range(m.n_elem):
        db.elems.append(Element(id=first_id + e, type="HEX8_G2" if m.elem_type[e] == HEX8_G2 else "HEX8_G1", pid=int(m.elem_pid[e]),
                                matid=int(m.elem_mat[e]) + 1, nlist=[int(v) + first_id for v in m.conn[e]]))
    for k, (E, nu) in enumerate(zip(m.mat_E, m.mat_nu)):
        db.mats.append(Material(id=k + 1, type="Elastic", name=f"Mat{k + 1}", E=float(E), poisson=float(nu), colorid=(k + 1) % 9))
    spc = BoundaryCondition(type="SPC", name="Fix", id=1, colorid=1,
                            nodal=[(int(n) + first_id, MatrixST(list(map(float, v)), 3, 1)) for n, v in zip(m.spc_node, m.spc_val)])
    load = BoundaryCondition(type="PointLoad", name="Load", id=2, colorid=2,
                             nodal=[(int(n) + first_id, MatrixST(list(map(float, v)), 3, 1)) for n, v in zip(m.load_node, m.load_val)])
    db.bcs = [(1, spc), (2, load)]
    db.ndof = m.n_dof
    db.analysis = Analysis(linsolver=m.lin_solver, tolerance=float(m.tolerance), itermax=int(m.max_iter), incnumb=1)
    return db


def results(db: Database):
    """(node_index, disp (n,3), strain (ne,8,6), stress (ne,8,6)) of increment 1."""
    ni = np.array([n.dof[0] // 3 for n in db.nodes], dtype=np.int32)
    disp = np.array([[n.dispx[1], n.dispy[1], n.dispz[1]] for n in db.nodes])
    strain = np.array([e.strain[1].array() for e in db.elems])
    stress = np.array([e.stress[1].array() for e in db.elems])
    return ni, disp, strain, stress


def to_model(db: Database) -> Model:
    """The flat model Solver.SolverLinearStatics reads from a database (Solver.cs:71-152): positions in
    NodeLib / ElemLib / MatLib order replace IDs, SPC and PointLoad entries are concatenated in BCLib order.
    Raises on what the native path does not cover (non-hex8 elements, unknown node or material IDs)."""
    node_pos = {n.id: i for i, n in enumerate(db.nodes)}
    mat_pos = {m.id: i for i, m in enumerate(db.mats)}
    types = {"HEX8_G2": HEX8_G2, "HEX8_G1": HEX8_G1}
    for e in db.elems:
        if len(e.nlist) != 8:
            raise ValueError(f"element {e.id} has {len(e.nlist)} nodes; the linear-static path takes hex8 only")
    try:
        conn = np.array([[node_pos[v] for v in e.nlist] for e in db.elems], dtype=np.int32).reshape(-1, 8)
        etype = np.array([types[e.type] for e in db.elems], dtype=np.uint8)
        emat = np.array([mat_pos[e.matid] for e in db.elems], dtype=np.int32)
    except KeyError as k:
        raise ValueError(f"database entry {k} is outside the linear-static hex8 path") from None
    spc_n, spc_v, load_n, load_v = [], [], [], []
    for _, bc in db.bcs:
        dst_n, dst_v = (spc_n, spc_v) if bc.type == "SPC" else (load_n, load_v) if bc.type == "PointLoad" else (None, None)
        if dst_n is None:
            continue
        for nid, mat in bc.nodal:
            if nid not in node_pos:
                raise ValueError(f"BC references missing node {nid}")
            dst_n.append(node_pos[nid])
            dst_v.append((list(mat.M) + [0.0, 0.0, 0.0])[:3])
    a = db.analysis or Analysis()
    return Model(xyz=np.array([[n.x, n.y, n.z] for n in db.nodes], dtype=np.float64).reshape(-1, 3), conn=conn, elem_type=etype,
                 elem_mat=emat, elem_pid=np.array([e.pid for e in db.elems], dtype=np.int32),
                 mat_E=np.array([m.E for m in db.mats]), mat_nu=np.array([m.poisson for m in db.mats]),
                 spc_node=np.array(spc_n, dtype=np.int32), spc_val=np.array(spc_v, dtype=np.float64).reshape(-1, 3),
                 load_node=np.array(load_n, dtype=np.int32), load_val=np.array(load_v, dtype=np.float64).reshape(-1, 3),
                 tolerance=a.tolerance, max_iter=a.itermax, lin_solver=a.linsolver)
