"""Synthetic structured hex8 meshes for the linear-static hot path.

Flat model = the flattening of the reference's object graph that the C ABI consumes
(include/stan_b200.h): arrays in NodeLib / ElemLib insertion order, 0-based node indices.
CHEXA local node order follows /root/reference/src/STAN_Database/FE_Library.cs:108-115.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

HEX8_G1 = 1
HEX8_G2 = 2

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64_next(state: np.ndarray):
    """One splitmix64 step on a vector of uint64 states: returns (new_state, output)."""
    with np.errstate(over="ignore"):
        state = (state + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = state.copy()
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        z = z ^ (z >> np.uint64(31))
    return state, z


@dataclass
class Model:
    """Flat linear-static model.  Mirrors what Solver.SolverLinearStatics reads from Database
    (/root/reference/src/STAN_Solver/Solver.cs:71-152)."""

    xyz: np.ndarray            # (n_nodes, 3) float64
    conn: np.ndarray           # (n_elem, 8) int32, 0-based node indices
    elem_type: np.ndarray      # (n_elem,) uint8: 1 = HEX8_G1, 2 = HEX8_G2
    elem_mat: np.ndarray       # (n_elem,) int32 index into mat_E / mat_nu
    elem_pid: np.ndarray       # (n_elem,) int32 part id (1-based, BDF field 3)
    mat_E: np.ndarray          # (n_mat,) float64
    mat_nu: np.ndarray         # (n_mat,) float64
    spc_node: np.ndarray       # (n_spc,) int32
    spc_val: np.ndarray        # (n_spc, 3) float64, 1 = fixed (Solver.cs:110-112)
    load_node: np.ndarray      # (n_load,) int32
    load_val: np.ndarray       # (n_load, 3) float64
    tolerance: float = 1.0e-8  # Analysis.LinSolverTolerance
    max_iter: int = 0          # Analysis.LinSolverIterMax
    lin_solver: str = "CG"     # Analysis.LinSolver: "CG" | "Cholesky" (Solver.cs:162-164)
    dims: tuple = field(default=(0, 0, 0))

    @property
    def n_nodes(self) -> int:
        return int(self.xyz.shape[0])

    @property
    def n_elem(self) -> int:
        return int(self.conn.shape[0])

    @property
    def n_dof(self) -> int:
        return 3 * self.n_nodes


def beam(nx: int, ny: int, nz: int, *, h: float = 1.0, elem_type: int = HEX8_G2, jitter: bool = False,
         seed: int = 12345, n_parts: int = 1, E=(210000.0, 70000.0), nu=(0.3, 0.33),
         total_load: float = 1000.0, tolerance: float = 1.0e-8, max_iter: int = 0) -> Model:
    """nx x ny x nz block of cubes (beam axis z), clamped at k = 0, tip load Fx at k = nz.

    node id = 1 + i + (nx+1)(j + (ny+1)k) (x fastest), index = id - 1; elements likewise.
    jitter: strictly interior nodes move by U(-0.1h, 0.1h) per coordinate, three successive
    splitmix64 outputs of state (seed + node_id) — removes exact-zero couplings so the CSR
    pattern is unambiguous (SURVEY.md §7 "Pattern definition").
    n_parts: slabs along z with materials alternating E[0]/nu[0], E[1]/nu[1].
    """
    nxn, nyn, nzn = nx + 1, ny + 1, nz + 1
    k, j, i = np.meshgrid(np.arange(nzn), np.arange(nyn), np.arange(nxn), indexing="ij")
    xyz = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64) * h
    n_nodes = xyz.shape[0]
    if jitter:
        interior = ((i > 0) & (i < nx) & (j > 0) & (j < ny) & (k > 0) & (k < nz)).ravel()
        state = (np.arange(1, n_nodes + 1, dtype=np.uint64) + np.uint64(seed)) & _M64
        for c in range(3):
            state, out = _splitmix64_next(state)
            u = (out >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
            xyz[:, c] += np.where(interior, (2.0 * u - 1.0) * 0.1 * h, 0.0)

    ek, ej, ei = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (ei + nxn * (ej + nyn * ek)).ravel().astype(np.int64)
    off = np.array([0, 1, 1 + nxn, nxn, nxn * nyn, nxn * nyn + 1, nxn * nyn + 1 + nxn, nxn * nyn + nxn],
                   dtype=np.int64)
    conn = (base[:, None] + off[None, :]).astype(np.int32)
    n_elem = conn.shape[0]

    part = np.minimum((ek.ravel() * n_parts) // max(nz, 1), n_parts - 1).astype(np.int32)
    n_mat = 1 if n_parts == 1 else 2
    elem_mat = (part % n_mat).astype(np.int32)
    spc_node = np.arange(nxn * nyn, dtype=np.int32)                      # k = 0 face
    spc_val = np.ones((spc_node.size, 3), dtype=np.float64)
    load_node = (np.arange(nxn * nyn, dtype=np.int64) + nxn * nyn * nz).astype(np.int32)
    load_val = np.zeros((load_node.size, 3), dtype=np.float64)
    load_val[:, 0] = total_load / load_node.size
    return Model(xyz=np.ascontiguousarray(xyz), conn=np.ascontiguousarray(conn),
                 elem_type=np.full(n_elem, elem_type, dtype=np.uint8), elem_mat=elem_mat,
                 elem_pid=(part + 1).astype(np.int32),
                 mat_E=np.asarray(E[:n_mat], dtype=np.float64), mat_nu=np.asarray(nu[:n_mat], dtype=np.float64),
                 spc_node=spc_node, spc_val=spc_val, load_node=load_node, load_val=load_val,
                 tolerance=tolerance, max_iter=max_iter, dims=(nx, ny, nz))


# Named workloads of BASELINE.json / SURVEY.md §8.
WORKLOADS = {
    "beam_100k_g2": dict(nx=20, ny=20, nz=250, elem_type=HEX8_G2),
    "beam_1m_g1": dict(nx=49, ny=51, nz=400, elem_type=HEX8_G1),
    "beam_10m_g2": dict(nx=100, ny=100, nz=1000, elem_type=HEX8_G2),
    "block_40m_g2": dict(nx=400, ny=400, nz=250, elem_type=HEX8_G2, n_parts=4),
}


def workload(name: str, **overrides) -> Model:
    kw = dict(WORKLOADS[name])
    kw.update(overrides)
    return beam(**kw)


def write_bdf(model: Model, path: str) -> None:
    """Short-format Nastran deck as the reference imports it
    (/root/reference/README.md:35-48; parsers Node.cs:25-80, Element.cs:35-73)."""
    def f8(v: float) -> str:
        s = f"{v:.6g}"
        if "e" in s or "E" in s or len(s) > 8:
            s = f"{v:8.5f}"[:8]
        if "." not in s:
            s += "."
        return s.rjust(8)

    with open(path, "w") as fh:
        fh.write("$$  GRID Data\n")
        for n, (x, y, z) in enumerate(model.xyz, start=1):
            fh.write(f"GRID    {n:8d}        {f8(x)}{f8(y)}{f8(z)}\n")
        fh.write("$$  CHEXA Elements: First Order\n")
        for e, (nodes, pid) in enumerate(zip(model.conn + 1, model.elem_pid), start=1):
            a = "".join(f"{int(v):8d}" for v in nodes[:6])
            b = "".join(f"{int(v):8d}" for v in nodes[6:])
            fh.write(f"CHEXA   {e:8d}{int(pid):8d}{a}+\n+       {b}\n")


def polar_disk(n_sectors: int, n_rings: int, nz: int, *, r_outer: float = 10.0, height: float = 5.0,
               elem_type: int = HEX8_G2, E: float = 210000.0, nu: float = 0.3, tolerance: float = 1.0e-8) -> Model:
    """Unstructured-valence test mesh: a disk meshed in polar fashion, extruded in z.

    The axis nodes touch 2*n_sectors elements (far more than the 8 of a structured grid) and the
    innermost ring consists of hexahedra collapsed to wedges — the axis node appears twice in their
    connectivity — so rows with many blocks, many incident elements and degenerate elements are all
    exercised.  Bottom face (z = 0) clamped, outer rim of the top face loaded in +x.
    """
    def nid(k, s, l):                         # ring k (0 = axis), sector s, layer l
        if k == 0:
            return l
        return (nz + 1) + ((l * n_rings) + (k - 1)) * n_sectors + (s % n_sectors)

    n_nodes = (nz + 1) * (1 + n_rings * n_sectors)
    xyz = np.zeros((n_nodes, 3))
    for l in range(nz + 1):
        z = height * l / nz
        xyz[nid(0, 0, l)] = (0.0, 0.0, z)
        for k in range(1, n_rings + 1):
            r = r_outer * k / n_rings
            for s in range(n_sectors):
                t = 2.0 * np.pi * s / n_sectors
                xyz[nid(k, s, l)] = (r * np.cos(t), r * np.sin(t), z)
    conn = []
    for l in range(nz):
        for k in range(n_rings):
            for s in range(n_sectors):
                conn.append([nid(k, s, l), nid(k + 1, s, l), nid(k + 1, s + 1, l), nid(k, s + 1, l),
                             nid(k, s, l + 1), nid(k + 1, s, l + 1), nid(k + 1, s + 1, l + 1), nid(k, s + 1, l + 1)])
    conn = np.asarray(conn, dtype=np.int32)
    n_elem = conn.shape[0]
    bottom = np.array(sorted({nid(k, s, 0) for k in range(n_rings + 1) for s in range(n_sectors)}), dtype=np.int32)
    rim = np.array([nid(n_rings, s, nz) for s in range(n_sectors)], dtype=np.int32)
    load_val = np.zeros((rim.size, 3))
    load_val[:, 0] = 100.0 / rim.size
    return Model(xyz=xyz, conn=conn, elem_type=np.full(n_elem, elem_type, dtype=np.uint8),
                 elem_mat=np.zeros(n_elem, dtype=np.int32), elem_pid=np.ones(n_elem, dtype=np.int32),
                 mat_E=np.array([E]), mat_nu=np.array([nu]), spc_node=bottom, spc_val=np.ones((bottom.size, 3)),
                 load_node=rim, load_val=load_val, tolerance=tolerance, dims=(n_sectors, n_rings, nz))
