"""Host-side mirror of the reference's solver interface for the linear-static path.

Method names and argument meaning follow the managed code this package replaces so the parity
tests read like the reference: Database.AssignDOF (Database.cs:140-234),
SolverFunctions.ParallelAssembly_K / LinearSolver_CG / Include_BC_DOF / Exclude_BC_DOF
(SolverFunctions.cs:117,270,520,540), Element.Recovery_Stress (Element.cs:211) and the driver
Solver.SolverLinearStatics (Solver.cs:71-217).  Everything numeric happens in libstan_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import native
from .mesh import Model


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


@dataclass
class LinearStaticsResult:
    node_index: np.ndarray        # Node.DOF[0] // 3
    U_full: np.ndarray            # DOF order, zeros at fixed DOFs (Include_BC_DOF)
    disp: np.ndarray              # (n_nodes, 3): Node.DispX/Y/Z[1]
    strain: np.ndarray            # (n_elem, 8, 6): Element.Strain[1]
    stress: np.ndarray            # (n_elem, 8, 6): Element.Stress[1]
    assembly: native.AssemblyStats
    cg: native.CgReport
    recovery: native.RecoveryStats


class Solver:
    """One handle = one GPU.  Not re-entrant (the reference's solver is single-threaded too)."""

    def __init__(self, device: int = -1, rank: int = 0, world: int = 1, pinned_results: bool = False):
        """pinned_results: result arrays (U, displacements, strain, stress) live in page-locked buffers owned by
        this Solver and are REUSED by the next call — the device-to-host copies become plain DMA transfers."""
        self._lib = native.load()
        self._h = C.c_void_p()
        opts = native.Options(device, rank, world, 0)
        native.check(self._lib.stan_create(C.byref(opts), C.byref(self._h)))
        self.model: Model | None = None
        self.node_index: np.ndarray | None = None
        self.world = world
        self._pinned_results = pinned_results
        self._pinned = []                # (pointer, array) pairs from stan_host_alloc, freed by close()
        self._out = {}                   # name -> persistent result buffer

    def close(self):
        if self._h:
            self._out.clear()
            for ptr, _ in self._pinned:
                self._lib.stan_host_free(ptr)
            self._pinned.clear()
            self._lib.stan_destroy(self._h)
            self._h = C.c_void_p()

    def pinned_empty(self, shape, dtype=np.float64) -> np.ndarray:
        """numpy array in page-locked host memory (stan_host_alloc); valid until close()."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p()
        native.check(self._lib.stan_host_alloc(max(nbytes, 1), C.byref(ptr)))
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned.append((ptr, arr))
        return arr

    def pinned_model(self, m: Model) -> Model:
        """Copy of the model whose arrays are page-locked: SetModel/SetDOF then upload by DMA without staging."""
        import dataclasses
        kw = {}
        for f in dataclasses.fields(m):
            v = getattr(m, f.name)
            if isinstance(v, np.ndarray):
                a = self.pinned_empty(v.shape, v.dtype)
                a[...] = v
                v = a
            kw[f.name] = v
        return Model(**kw)

    def _result(self, name, shape, dtype=np.float64):
        if not self._pinned_results:
            return np.empty(shape, dtype=dtype)
        a = self._out.get(name)
        if a is None or a.shape != tuple(shape) or a.dtype != np.dtype(dtype):
            a = self._out[name] = self.pinned_empty(shape, dtype)
        return a

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- model upload (Solver.cs:26-39, 81-152 read these from the Database) ----
    def SetModel(self, m: Model):
        self.model = m
        xyz, conn = _c(m.xyz, np.float64), _c(m.conn, np.int32)
        et, em = _c(m.elem_type, np.uint8), _c(m.elem_mat, np.int32)
        native.check(self._lib.stan_set_mesh(self._h, m.n_nodes, _p(xyz), m.n_elem, _p(conn), _p(et), _p(em)))
        E, nu = _c(m.mat_E, np.float64), _c(m.mat_nu, np.float64)
        native.check(self._lib.stan_set_materials(self._h, len(E), _p(E), _p(nu)))
        sn, sv = _c(m.spc_node, np.int32), _c(m.spc_val, np.float64)
        native.check(self._lib.stan_set_spc(self._h, len(sn), _p(sn), _p(sv)))
        ln, lv = _c(m.load_node, np.int32), _c(m.load_val, np.float64)
        native.check(self._lib.stan_set_loads(self._h, len(ln), _p(ln), _p(lv)))
        self.node_index = None

    def AssignDOF(self) -> np.ndarray:
        """Database.AssignDOF: BFS index per node; DOF = 3*index + {0,1,2}."""
        out = np.zeros(self.model.n_nodes, dtype=np.int32)
        native.check(self._lib.stan_assign_dof(self._h, _p(out)))
        self.node_index = out
        return out

    def SetDOF(self, node_index):
        ni = _c(node_index, np.int32)
        native.check(self._lib.stan_set_dof_map(self._h, _p(ni)))
        self.node_index = ni

    # ---- hot path ----
    def ParallelAssembly_K(self) -> native.AssemblyStats:
        st = native.AssemblyStats()
        native.check(self._lib.stan_assemble(self._h, C.byref(st)))
        return st

    def LinearSolver_CG(self, tolerance=None, IterMax=None, *, merit_check=1, its_before_rupdate=10,
                        its_before_restart=0, zero_based_counter=0, time_kernels=0) -> native.CgReport:
        m = self.model
        o = native.CgOptions(m.tolerance if tolerance is None else tolerance, m.max_iter if IterMax is None else IterMax,
                             its_before_rupdate, its_before_restart, merit_check, zero_based_counter, time_kernels, 0)
        rep = native.CgReport()
        native.check(self._lib.stan_solve_cg(self._h, C.byref(o), C.byref(rep)))
        return rep

    def LinearSolver_Cholesky(self) -> native.CholReport:
        """Fun.LinearSolver_Cholesky(K, F) (SolverFunctions.cs:332-444): skyline U^T U on one GPU."""
        rep = native.CholReport()
        native.check(self._lib.stan_solve_cholesky(self._h, C.byref(rep)))
        return rep

    def Recovery_Stress(self) -> native.RecoveryStats:
        st = native.RecoveryStats()
        native.check(self._lib.stan_recover(self._h, C.byref(st)))
        return st

    # ---- results / inspection ----
    def Include_BC_DOF(self) -> np.ndarray:
        """U_Full (SolverFunctions.cs:520-538): displacement per DOF with zeros at SPC DOFs."""
        u = self._result("U", (self.model.n_dof,))
        native.check(self._lib.stan_get_displacements(self._h, _p(u)))
        return u

    def displacements_local(self) -> np.ndarray:
        """U of the rows this rank owns ((last_row - first_row) x 3); all of U on one GPU."""
        r0, r1 = self.partition()
        u = self._result("U_local", (r1 - r0, 3))
        native.check(self._lib.stan_get_displacements_local(self._h, _p(u)))
        return u

    def node_displacements(self) -> np.ndarray:
        """Node.dU_buffer per node in NodeLib order (Solver.cs:171-178), gathered through the DOF map on the device."""
        d = self._result("disp", (self.model.n_nodes, 3))
        native.check(self._lib.stan_get_node_displacements(self._h, _p(d)))
        return d

    def Exclude_BC_DOF(self) -> np.ndarray:
        n, _ = self.csr_upper_size()
        u = np.zeros(n)
        native.check(self._lib.stan_get_solution_reduced(self._h, _p(u)))
        return u

    def element_range(self):
        a, b = C.c_int64(), C.c_int64()
        native.check(self._lib.stan_get_element_range(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def strain_stress(self):
        """Strain[1]/Stress[1] of this rank's element slice (all elements on one GPU)."""
        e0, e1 = self.element_range()
        ne = e1 - e0
        strain, stress = self._result("strain", (ne, 8, 6)), self._result("stress", (ne, 8, 6))
        native.check(self._lib.stan_get_strain_stress(self._h, _p(strain), _p(stress)))
        return strain, stress

    def Load_Scalar(self):
        """Part.Load_Scalar (Part.cs:231-528): returns (cell (n_elem,24,3) max/average/min, point (n_nodes,24), ms)."""
        ms = C.c_double()
        native.check(self._lib.stan_postprocess(self._h, C.byref(ms)))
        e0, e1 = self.element_range()
        r0, r1 = self.partition()
        cell = np.empty((e1 - e0, 24, 3), dtype=np.float32)
        point = np.empty((r1 - r0, 24), dtype=np.float32)      # NodeLib order on one GPU, row order when partitioned
        native.check(self._lib.stan_get_scalars(self._h, _p(cell), _p(point)))
        return cell, point, ms.value

    def nDOF_reduction(self) -> np.ndarray:
        red = np.zeros(self.model.n_dof, dtype=np.int32)
        native.check(self._lib.stan_get_dof_reduction(self._h, _p(red)))
        return red

    def F(self) -> np.ndarray:
        n, _ = self.csr_upper_size()
        f = np.zeros(n)
        native.check(self._lib.stan_get_rhs(self._h, _p(f)))
        return f

    def csr_upper_size(self):
        n, nnz = C.c_int64(), C.c_int64()
        native.check(self._lib.stan_get_csr_upper_size(self._h, C.byref(n), C.byref(nnz)))
        return n.value, nnz.value

    def csr_upper(self):
        n, nnz = self.csr_upper_size()
        rp, col, val = np.zeros(n + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz)
        native.check(self._lib.stan_get_csr_upper(self._h, _p(rp), _p(col), _p(val)))
        return rp, col, val

    def K_Initial(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.model.n_elem - first if count is None else count
        ke = np.zeros((count, 24, 24))
        native.check(self._lib.stan_element_stiffness(self._h, first, count, _p(ke)))
        return ke

    def spmv(self, x_full) -> np.ndarray:
        x = _c(x_full, np.float64)
        y = np.zeros_like(x)
        native.check(self._lib.stan_spmv(self._h, _p(x), _p(y)))
        return y

    def cg_history(self, capacity: int | None = None):
        """capacity given: arm the recorder for the next LinearSolver_CG; else fetch the (count, 4) record of
        the last solve: ||r_k||^2, alpha_k, beta_k, energy functional on refresh iterations (NaN otherwise)."""
        if capacity is not None:
            native.check(self._lib.stan_set_cg_history(self._h, int(capacity)))
            return None
        n = C.c_int32()
        native.check(self._lib.stan_get_cg_history(self._h, C.byref(n), None))
        hist = np.empty((n.value, 4))
        if n.value:
            native.check(self._lib.stan_get_cg_history(self._h, C.byref(n), _p(hist)))
        return hist

    def time_spmv(self, reps: int = 20):
        ms, by = C.c_double(), C.c_int64()
        native.check(self._lib.stan_time_spmv(self._h, reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def kernel_launches(self) -> int:
        return int(self._lib.stan_kernel_launches(self._h))

    def event_record(self, slot: int):
        native.check(self._lib.stan_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        native.check(self._lib.stan_event_elapsed(self._h, a, b, C.byref(ms)))
        return ms.value

    def partition(self):
        a, b = C.c_int64(), C.c_int64()
        native.check(self._lib.stan_get_partition(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        native.check(self._lib.stan_comm_init(self._h, buf))

    # ---- the driver: Solver.SolverLinearStatics (Solver.cs:71-217) ----
    def SolverLinearStatics(self, m: Model, *, node_index=None, merit_check=1, time_kernels=0,
                            fetch=True, local_rows=False) -> LinearStaticsResult:
        """local_rows (several GPUs): U_full holds only the rows this rank owns (partition() x 3) and disp is None —
        the caller merges the ranks' slices, as it already does for the strain/stress slices."""
        self.SetModel(m)
        if node_index is None:
            self.AssignDOF()
        else:
            self.SetDOF(node_index)
        a = self.ParallelAssembly_K()
        if m.lin_solver == "CG":                                   # Solver.cs:162-164
            cg = self.LinearSolver_CG(merit_check=merit_check, time_kernels=time_kernels)
        elif m.lin_solver == "Cholesky":
            cg = self.LinearSolver_Cholesky()
        else:
            raise ValueError(f"LinSolver {m.lin_solver!r} is not provided (CG and Cholesky are)")
        rec = self.Recovery_Stress()
        if not fetch:
            return LinearStaticsResult(self.node_index, None, None, None, None, a, cg, rec)
        strain, stress = self.strain_stress()
        if local_rows:
            return LinearStaticsResult(self.node_index, self.displacements_local(), None, strain, stress, a, cg, rec)
        U = self.Include_BC_DOF()
        disp = self.node_displacements()
        return LinearStaticsResult(self.node_index, U, disp, strain, stress, a, cg, rec)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    native.check(native.load().stan_comm_unique_id(buf))
    return buf.raw
