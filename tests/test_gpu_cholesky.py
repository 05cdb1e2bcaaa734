"""LinearSolver_Cholesky on the GPU (SolverFunctions.cs:332-444) against the oracle's skyline U^T U.

A direct solve has no iteration history to match; parity is the solution itself.  Both sides
factorise the same matrix in IEEE double with different (fixed) summation orders, so they agree to
a small multiple of cond(K)*eps: 1e-10 relative on these meshes, asserted below, and the stresses
recovered from it to 1e-8 (the bars BASELINE.json states for the path)."""
import numpy as np
import pytest

from stan_b200 import mesh, native
from stan_b200.solver import Solver

pytestmark = pytest.mark.gpu


def _oracle_solution(oracle, m, ni):
    red, _ = oracle.spc_reduction(m, ni)
    K = oracle.assemble_upper(m, ni, red)
    b = oracle.build_rhs(m, ni, red)
    x, tt, env = oracle.cholesky_skyline(K, b)
    return red, x, tt, env


@pytest.mark.parametrize("dims,jit", [((1, 1, 1), False), ((3, 3, 8), True), ((6, 5, 20), True), ((7, 2, 9), False),
                                      ((10, 10, 30), True)])
def test_cholesky_matches_the_oracle(oracle, dims, jit):
    # HEX8_G2 only: one-point integration (HEX8_G1) leaves hourglass modes, K is singular and both
    # sides report -3 (covered below by the not-positive-definite case).
    m = mesh.beam(*dims, elem_type=mesh.HEX8_G2, jitter=jit)
    with Solver() as s:
        s.SetModel(m); ni = s.AssignDOF(); s.ParallelAssembly_K()
        rep = s.LinearSolver_Cholesky()
        red, xo, tt, env = _oracle_solution(oracle, m, ni)
        assert (rep.terminationtype, tt) == (1, 1)
        assert rep.n == m.n_dof and rep.block == 64
        assert rep.skyline_bytes == rep.n_blocks * 64 * 64 * 8 >= env * 8      # the block skyline covers the envelope
        x = s.Exclude_BC_DOF()
        assert np.abs(x - xo).max() <= 1e-10 * np.abs(xo).max()
        U = s.Include_BC_DOF()
        assert np.all(U[red == -1] == 0.0)
        s.Recovery_Stress()
        strain, stress = s.strain_stress()
        so, to = oracle.recover(m, ni, oracle.include_bc_dof(red, xo))
        assert np.abs(stress - to).max() <= 1e-8 * np.abs(to).max()
        assert np.abs(strain - so).max() <= 1e-8 * np.abs(so).max()


def test_cholesky_agrees_with_cg_and_is_reproducible(oracle):
    m = mesh.beam(8, 7, 40, jitter=True, tolerance=1e-10)
    with Solver() as s:
        s.SetModel(m); s.AssignDOF(); s.ParallelAssembly_K()
        s.LinearSolver_Cholesky()
        x1 = s.Exclude_BC_DOF()
        s.LinearSolver_Cholesky()
        x2 = s.Exclude_BC_DOF()
        assert np.array_equal(x1, x2)                         # fixed summation order: bit-identical reruns
        cg = s.LinearSolver_CG(merit_check=0, IterMax=20000)
        assert cg.terminationtype == 1
        xc = s.Exclude_BC_DOF()
        assert np.linalg.norm(xc - x1) <= 1e-7 * np.linalg.norm(x1)
        # residual of the direct solution through the product's own SpMV
        U = np.zeros(m.n_dof); red = s.nDOF_reduction(); U[red >= 0] = x1
        r = s.spmv(U)[red >= 0] - s.F()
        assert np.linalg.norm(r) <= 1e-10 * np.linalg.norm(s.F())


def test_panel_grid_larger_than_one_wave(monkeypatch):
    """A block row with more blocks than one resident wave of k_chol_panel CTAs (the later CTAs used to
    read a diagonal block that CTA 0 had already overwritten with its factor).  56x56 cross-section: half-bandwidth
    ~9.9 k DOF = 155 blocks; with one block per CTA that is more than the 148 SMs hold at once.  The factorisation
    is the same arithmetic whatever the grid shape, so the two runs must agree bit for bit, and with CG."""
    m = mesh.beam(56, 56, 6, tolerance=1e-10)
    with Solver() as s:
        s.SetModel(m); s.AssignDOF(); s.ParallelAssembly_K()
        rep = s.LinearSolver_Cholesky()
        assert rep.terminationtype == 1
        x_default = s.Exclude_BC_DOF()
        monkeypatch.setenv("STAN_CHOL_TILES", "1")
        rep1 = s.LinearSolver_Cholesky()
        assert rep1.terminationtype == 1 and rep1.n_blocks == rep.n_blocks
        x_one = s.Exclude_BC_DOF()
        assert np.array_equal(x_one, x_default)
        U = np.zeros(m.n_dof); red = s.nDOF_reduction(); U[red >= 0] = x_one
        r = s.spmv(U)[red >= 0] - s.F()
        assert np.linalg.norm(r) <= 1e-9 * np.linalg.norm(s.F())
        monkeypatch.delenv("STAN_CHOL_TILES")
        cg = s.LinearSolver_CG(merit_check=0, IterMax=20000)
        assert cg.terminationtype == 1
        assert np.linalg.norm(s.Exclude_BC_DOF() - x_one) <= 1e-7 * np.linalg.norm(x_one)


def test_not_positive_definite_reports_minus_3_and_zeros(oracle):
    m = mesh.beam(3, 3, 6)
    m.spc_node = m.spc_node[:0]; m.spc_val = m.spc_val[:0]   # rigid-body modes: no Cholesky factor
    m.mat_E = -m.mat_E                                        # and make it negative definite for a clean failure
    with Solver() as s:
        s.SetModel(m); ni = s.AssignDOF(); s.ParallelAssembly_K()
        rep = s.LinearSolver_Cholesky()
        _, xo, tt, _ = _oracle_solution(oracle, m, ni)
        assert rep.terminationtype == -3 and tt == -3
        assert not s.Include_BC_DOF().any() and not xo.any()  # "filled by zeros" (SolverFunctions.cs:420)


def test_driver_dispatches_on_lin_solver(oracle):
    m = mesh.beam(4, 4, 10, jitter=True)
    m.lin_solver = "Cholesky"
    with Solver() as s:
        r = s.SolverLinearStatics(m)
        assert isinstance(r.cg, native.CholReport) and r.cg.terminationtype == 1
        o = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-10, maxits=5000, merit_check=0))
        assert np.linalg.norm(r.U_full - o.U_full) <= 1e-8 * np.linalg.norm(o.U_full)
        m.lin_solver = "LU"
        with pytest.raises(ValueError):
            s.SolverLinearStatics(m)


def test_skyline_that_cannot_fit_is_refused():
    m = mesh.beam(100, 100, 100)                              # 3.1M DOF, BFS bandwidth ~1e5: a multi-TB skyline
    with Solver() as s:
        s.SetModel(m); s.AssignDOF(); s.ParallelAssembly_K()
        with pytest.raises(native.StanError) as ei:
            s.LinearSolver_Cholesky()
        assert ei.value.code == native.E_NOMEM and "use CG" in str(ei.value)
        assert s.LinearSolver_CG(IterMax=5).iterationscount == 5      # the handle is still usable
