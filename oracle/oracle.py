"""ctypes front-end of the CPU oracle (oracle/stan_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the stan_b200 package.
PARITY UNPINNED (see the header of stan_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstan_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "stan_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class CgOpts(C.Structure):
    _fields_ = [("epsf", C.c_double), ("maxits", C.c_int32), ("its_before_rupdate", C.c_int32),
                ("its_before_restart", C.c_int32), ("merit_check", C.c_int32),
                ("zero_based_counter", C.c_int32), ("parallel_spmv", C.c_int32), ("dot_mode", C.c_int32)]


class CgReport(C.Structure):
    _fields_ = [("terminationtype", C.c_int32), ("iterationscount", C.c_int32), ("nmv", C.c_int32),
                ("reserved", C.c_int32), ("r2", C.c_double), ("bnorm", C.c_double)]


class PathStats(C.Structure):
    _fields_ = [("t_assign_dof", C.c_double), ("t_assembly", C.c_double), ("t_solve", C.c_double),
                ("t_recovery", C.c_double), ("t_total", C.c_double), ("n_free", C.c_int64),
                ("nnz_upper", C.c_int64), ("threads", C.c_int32), ("pad", C.c_int32), ("cg", CgReport)]


def cg_opts(epsf=1e-8, maxits=0, its_before_rupdate=10, its_before_restart=0, merit_check=1,
            zero_based_counter=0, parallel_spmv=0, dot_mode=0) -> CgOpts:
    return CgOpts(epsf, maxits, its_before_rupdate, its_before_restart, merit_check, zero_based_counter,
                  parallel_spmv, dot_mode)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.stan_oracle_assemble_upper.restype = C.c_void_p
        _lib.stan_oracle_csr_n.restype = C.c_int64
        _lib.stan_oracle_csr_nnz.restype = C.c_int64
        _lib.stan_oracle_csr_n.argtypes = [C.c_void_p]
        _lib.stan_oracle_csr_nnz.argtypes = [C.c_void_p]
        _lib.stan_oracle_csr_free.argtypes = [C.c_void_p]
        _lib.stan_oracle_csr_copy.argtypes = [C.c_void_p] * 4
        _lib.stan_oracle_elastic_D.argtypes = [C.c_double, C.c_double, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def hex8_tables(elem_type: int):
    ngp, w = C.c_int(), C.c_double()
    dN = np.zeros(8 * 24)
    N = np.zeros(64)
    rc = lib().stan_oracle_hex8_tables(int(elem_type), C.byref(ngp), C.byref(w), _p(dN), _p(N))
    assert rc == 0
    g = ngp.value
    return g, w.value, dN[: g * 24].reshape(g, 3, 8).copy(), (N.reshape(8, 8) if g == 8 else N[:8].reshape(1, 8)).copy()


def elastic_D(E: float, nu: float) -> np.ndarray:
    D = np.zeros(36)
    lib().stan_oracle_elastic_D(float(E), float(nu), _p(D))
    return D.reshape(6, 6)


def k_initial(elem_type: int, X, D, want_cache=False):
    X = _f64(X).reshape(24)
    D = _f64(D).reshape(36)
    K = np.zeros(576)
    ngp = 8 if elem_type == 2 else 1
    J = np.zeros(ngp * 9)
    BL = np.zeros(ngp * 144)
    rc = lib().stan_oracle_k_initial(int(elem_type), _p(X), _p(D), _p(K), _p(J), _p(BL))
    if rc:
        raise ArithmeticError(f"k_initial rc={rc}")
    if want_cache:
        return K.reshape(24, 24), J.reshape(ngp, 3, 3), BL.reshape(ngp, 6, 24)
    return K.reshape(24, 24)


def assign_dof(model) -> np.ndarray:
    out = np.zeros(model.n_nodes, dtype=np.int32)
    conn = _i32(model.conn)
    rc = lib().stan_oracle_assign_dof(model.n_nodes, model.n_elem, _p(conn), _p(out))
    if rc:
        raise RuntimeError(f"assign_dof rc={rc}")
    return out


def spc_reduction(model, node_index):
    red = np.zeros(model.n_dof, dtype=np.int32)
    ni, sn, sv = _i32(node_index), _i32(model.spc_node), _f64(model.spc_val)
    nfix = lib().stan_oracle_spc_reduction(model.n_nodes, _p(ni), len(sn), _p(sn), _p(sv), _p(red))
    return red, nfix


def build_rhs(model, node_index, red):
    nfree = model.n_dof - int((red == -1).sum())
    F = np.zeros(nfree)
    ni, ln, lv = _i32(node_index), _i32(model.load_node), _f64(model.load_val)
    lib().stan_oracle_build_rhs(model.n_nodes, _p(ni), _p(red), len(ln), _p(ln), _p(lv), _p(F))
    return F


def include_bc_dof(red, A):
    out = np.zeros(len(red))
    A = _f64(A)
    lib().stan_oracle_include_bc_dof(len(red), _p(red), _p(A), _p(out))
    return out


class UpperCsr:
    """Owns an oracle_csr*; exposes numpy copies."""

    def __init__(self, handle, exact_zero):
        self._h = handle
        self.exact_zero = exact_zero
        self.n = lib().stan_oracle_csr_n(handle)
        self.nnz = lib().stan_oracle_csr_nnz(handle)
        self._arrays = None

    @classmethod
    def from_dense_upper(cls, A):
        """Upper triangle (nonzeros, diagonal always) of a small dense symmetric matrix — hand-worked test cases."""
        A = np.asarray(A, dtype=np.float64)
        n = A.shape[0]
        rp, col, val = [0], [], []
        for i in range(n):
            for j in range(i, n):
                if j == i or A[i, j] != 0.0:
                    col.append(j); val.append(A[i, j])
            rp.append(len(col))
        rp, col, val = np.array(rp, np.int64), np.array(col, np.int32), np.array(val)
        fn = lib().stan_oracle_csr_from_arrays
        fn.restype = C.c_void_p
        return cls(fn(C.c_int64(n), _p(rp), _p(col), _p(val)), 0)

    @classmethod
    def from_arrays(cls, rowptr, col, val):
        """Wraps an upper-triangle CRS given as arrays (e.g. the one libstan_b200 exports) for lincg / sym_spmv."""
        rp, c, v = np.ascontiguousarray(rowptr, np.int64), np.ascontiguousarray(col, np.int32), _f64(val)
        fn = lib().stan_oracle_csr_from_arrays
        fn.restype = C.c_void_p
        return cls(fn(C.c_int64(len(rp) - 1), _p(rp), _p(c), _p(v)), 0)

    def arrays(self):
        if self._arrays is None:
            rp = np.zeros(self.n + 1, dtype=np.int64)
            col = np.zeros(self.nnz, dtype=np.int32)
            val = np.zeros(self.nnz)
            lib().stan_oracle_csr_copy(self._h, _p(rp), _p(col), _p(val))
            self._arrays = (rp, col, val)
        return self._arrays

    def to_scipy_full(self):
        import scipy.sparse as sp
        rp, col, val = self.arrays()
        U = sp.csr_matrix((val, col, rp), shape=(self.n, self.n))
        return (U + sp.triu(U, 1).T).tocsr()

    def __del__(self):
        if self._h:
            lib().stan_oracle_csr_free(self._h)
            self._h = None


def assemble_upper(model, node_index, red, prune=False) -> UpperCsr:
    xyz, conn = _f64(model.xyz), _i32(model.conn)
    et = np.ascontiguousarray(model.elem_type, dtype=np.uint8)
    em, E, nu, ni = _i32(model.elem_mat), _f64(model.mat_E), _f64(model.mat_nu), _i32(node_index)
    err, ez = C.c_int(0), C.c_int64(0)
    h = lib().stan_oracle_assemble_upper(model.n_nodes, _p(xyz), model.n_elem, _p(conn), _p(et), _p(em), len(E),
                                         _p(E), _p(nu), _p(ni), _p(red), int(bool(prune)), C.byref(ez), C.byref(err))
    if not h:
        raise ArithmeticError(f"assemble_upper err={err.value}")
    return UpperCsr(h, ez.value)


def lincg(K: UpperCsr, b, opts: CgOpts):
    if not opts.merit_check and opts.maxits <= 0:
        raise ValueError("merit_check=0 needs maxits > 0 (an unreachable EpsF would never terminate)")
    b = _f64(b)
    x = np.zeros(K.n)
    rep = CgReport()
    lib().stan_oracle_lincg(C.c_void_p(K._h), _p(b), C.byref(opts), _p(x), C.byref(rep))
    return x, rep


def lincg_history(K: UpperCsr, b, opts: CgOpts, cap: int):
    """lincg plus its per-iteration trajectory: rows k = 1.. of (||r_k||^2, alpha_k, beta_k, energy functional
    on refresh iterations else NaN) — the same record stan_get_cg_history returns."""
    if not opts.merit_check and opts.maxits <= 0:
        raise ValueError("merit_check=0 needs maxits > 0")
    b = _f64(b)
    x = np.zeros(K.n)
    rep = CgReport()
    hist = np.full((cap, 4), np.nan)
    lib().stan_oracle_lincg_hist(C.c_void_p(K._h), _p(b), C.byref(opts), _p(x), C.byref(rep), _p(hist), C.c_int64(cap))
    return x, rep, hist[: min(cap, rep.iterationscount)]


def cholesky_skyline(K: UpperCsr, b):
    """LinearSolver_Cholesky (SolverFunctions.cs:332-444): returns (x, terminationtype, envelope size)."""
    b = _f64(b)
    x = np.zeros(K.n)
    env = C.c_int64(0)
    fn = lib().stan_oracle_cholesky_skyline
    fn.restype = C.c_int
    tt = fn(C.c_void_p(K._h), _p(b), _p(x), C.byref(env))
    return x, int(tt), int(env.value)


def sym_spmv(K: UpperCsr, x):
    x = _f64(x)
    y = np.zeros(K.n)
    lib().stan_oracle_sym_spmv(C.c_void_p(K._h), _p(x), _p(y))
    return y


def recover(model, node_index, U_full):
    xyz, conn = _f64(model.xyz), _i32(model.conn)
    et = np.ascontiguousarray(model.elem_type, dtype=np.uint8)
    em, E, nu, ni, U = _i32(model.elem_mat), _f64(model.mat_E), _f64(model.mat_nu), _i32(node_index), _f64(U_full)
    strain = np.zeros((model.n_elem, 8, 6))
    stress = np.zeros((model.n_elem, 8, 6))
    rc = lib().stan_oracle_recover(model.n_nodes, _p(xyz), model.n_elem, _p(conn), _p(et), _p(em), len(E), _p(E),
                                   _p(nu), _p(ni), _p(U), _p(strain), _p(stress))
    if rc:
        raise ArithmeticError(f"recover rc={rc}")
    return strain, stress


@dataclass
class PathResult:
    node_index: np.ndarray
    U_full: np.ndarray
    strain: np.ndarray
    stress: np.ndarray
    stats: PathStats


def linear_statics(model, opts: CgOpts | None = None) -> PathResult:
    """Whole path in one native call with a wall-clock breakdown (CPU baseline leg)."""
    opts = opts or cg_opts(epsf=model.tolerance, maxits=model.max_iter)
    xyz, conn = _f64(model.xyz), _i32(model.conn)
    et = np.ascontiguousarray(model.elem_type, dtype=np.uint8)
    em, E, nu = _i32(model.elem_mat), _f64(model.mat_E), _f64(model.mat_nu)
    sn, sv, ln, lv = _i32(model.spc_node), _f64(model.spc_val), _i32(model.load_node), _f64(model.load_val)
    ni = np.zeros(model.n_nodes, dtype=np.int32)
    U = np.zeros(model.n_dof)
    strain = np.zeros((model.n_elem, 8, 6))
    stress = np.zeros((model.n_elem, 8, 6))
    st = PathStats()
    rc = lib().stan_oracle_linear_statics(model.n_nodes, _p(xyz), model.n_elem, _p(conn), _p(et), _p(em), len(E),
                                          _p(E), _p(nu), len(sn), _p(sn), _p(sv), len(ln), _p(ln), _p(lv),
                                          C.byref(opts), _p(ni), _p(U), _p(strain), _p(stress), C.byref(st))
    if rc:
        raise RuntimeError(f"oracle linear_statics rc={rc}")
    return PathResult(ni, U, strain, stress, st)


def threads() -> int:
    return int(lib().stan_oracle_threads())


def set_threads(n: int | None = None) -> int:
    """OpenMP team size of the oracle; default = the cores this process may run on (torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently turn the CPU baseline single-threaded)."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().stan_oracle_set_threads(int(n))
    return threads()
