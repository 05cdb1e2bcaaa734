"""GPU parity against the CPU oracle at the sizes BASELINE.json names (VERDICT r01 "What's missing" 2).

The reference path being matched is Solver.SolverLinearStatics (/root/reference/src/STAN_Solver/Solver.cs:97-210).
The oracle is a restatement (PARITY UNPINNED, see oracle/stan_oracle.c): "matches the oracle" below means
"matches the CPU restatement of STAN", not "matches a STAN binary".

Bars (BASELINE.json north_star): DOF numbering and CSR pattern bit-exact, displacements 1e-10 relative and
stresses 1e-8 relative at an identical CG tolerance.  The tolerance is EpsF = 1e-9: the tightest these beams
reach in FP64 (the true residual of the 100k beam stalls near 1e-8..1e-9 * ||b|| when asked for 1e-10; measured
with the oracle, tools/cg_trajectory.py).  U is far better determined than the residual suggests — the load excites
the softest modes, so two converged runs agree to ~1e-12 — which is what makes the 1e-10 bar testable.
"""
import numpy as np
import pytest

from stan_b200 import mesh
from stan_b200.solver import Solver

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    s = Solver()
    yield s
    s.close()


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_beam_100k_g2_full_oracle_run(solver, oracle):
    """BASELINE config 2 (20x20x250 G2, 332 073 DOF): the whole path on both sides, strict CG."""
    oracle.set_threads()
    m = mesh.workload("beam_100k_g2", tolerance=1e-9)
    solver.SetModel(m)
    ni = solver.AssignDOF()
    assert np.array_equal(ni, oracle.assign_dof(m))                       # R0 bit-exact at size
    solver.ParallelAssembly_K()
    red, nfix = oracle.spc_reduction(m, ni)
    assert np.array_equal(solver.nDOF_reduction(), red)                   # R1
    F = oracle.build_rhs(m, ni, red)
    assert np.array_equal(solver.F(), F)
    K = oracle.assemble_upper(m, ni, red)
    rp, col, val = solver.csr_upper()
    orp, ocol, oval = K.arrays()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)          # R3 pattern bit-exact (12.7 M entries)
    assert np.array_equal(val, oval)                                      # R2+R3 values: bit-exact
    rep = solver.LinearSolver_CG(merit_check=0, IterMax=5000)             # R5, strict
    xo, orep = oracle.lincg(K, F, oracle.cg_opts(epsf=1e-9, merit_check=0, maxits=5000, parallel_spmv=1))
    assert rep.terminationtype == 1 and orep.terminationtype == 1
    # the oracle's own roundings need 1195 .. 1417 iterations here (profiles/r02_cg_trajectory_100k.json)
    assert abs(rep.iterationscount - orep.iterationscount) <= 0.25 * orep.iterationscount
    xg = solver.Exclude_BC_DOF()
    # R5 alone: the oracle's CG on the very matrix the device assembled
    xs, srep = oracle.lincg(oracle.UpperCsr.from_arrays(rp, col, val), F,
                            oracle.cg_opts(epsf=1e-9, merit_check=0, maxits=5000, parallel_spmv=1))
    assert srep.terminationtype == 1 and _rel(xg, xs) < 1e-10
    # whole path: device-assembled against oracle-assembled matrix
    assert _rel(xg, xo) < 1e-10
    solver.Recovery_Stress()                                              # R4
    strain, stress = solver.strain_stress()
    es, ss = oracle.recover(m, ni, oracle.include_bc_dof(red, xo))
    assert np.abs(stress - ss).max() <= 1e-8 * np.abs(ss).max()
    assert np.abs(strain - es).max() <= 1e-8 * np.abs(es).max()


def test_beam_1m_g1_assembly_spmv_recovery_and_first_iterations(solver, oracle):
    """BASELINE config 3 (49x51x400 G1, 3 127 800 DOF).  One-point integration without hourglass control makes the
    system so ill-conditioned that neither side converges in a test's time, so the solve is compared
    over a fixed budget of 30 iterations — where rounding has not yet separated two valid trajectories (the
    oracle's own second rounding is 1e-11 away by then and 1e-7 away by iteration 60 on G1 beams) — and
    assembly, the symmetric product and recovery are compared in full."""
    oracle.set_threads()
    m = mesh.workload("beam_1m_g1", tolerance=1e-30)
    solver.SetModel(m)
    ni = solver.AssignDOF()
    assert np.array_equal(ni, oracle.assign_dof(m))
    solver.ParallelAssembly_K()
    red, nfix = oracle.spc_reduction(m, ni)
    K = oracle.assemble_upper(m, ni, red)
    rp, col, val = solver.csr_upper()
    orp, ocol, oval = K.arrays()
    assert np.array_equal(rp, orp) and np.array_equal(col, ocol)          # 127 M entries, bit-exact pattern
    assert np.array_equal(val, oval)                                      # and values
    del rp, col, val
    xr = np.random.default_rng(11).standard_normal(K.n)
    y = solver.spmv(oracle.include_bc_dof(red, xr))[red != -1]
    yo = oracle.sym_spmv(K, xr)
    assert np.abs(y - yo).max() <= 1e-12 * np.abs(yo).max()
    F = oracle.build_rhs(m, ni, red)
    solver.cg_history(64)
    rep = solver.LinearSolver_CG(merit_check=0, IterMax=30)
    xo, orep, ho = oracle.lincg_history(K, F, oracle.cg_opts(epsf=1e-30, merit_check=0, maxits=30, parallel_spmv=1), 64)
    hg = solver.cg_history()
    solver.cg_history(0)
    assert rep.terminationtype == 5 and orep.terminationtype == 5 and rep.iterationscount == orep.iterationscount == 30
    assert rep.nmv == orep.nmv == 1 + 30 + 3
    assert hg.shape == ho.shape == (30, 4)
    rel = np.abs(hg[:, :3] - ho[:, :3]) / np.maximum(np.abs(ho[:, :3]), 1e-300)
    assert rel.max() < 1e-9
    assert np.array_equal(np.isfinite(hg[:, 3]), np.isfinite(ho[:, 3]))   # refreshes at 10, 20, 30 on both sides
    xg = solver.Exclude_BC_DOF()
    assert _rel(xg, xo) < 1e-8
    solver.Recovery_Stress()
    U = solver.Include_BC_DOF()
    strain, stress = solver.strain_stress()
    es, ss = oracle.recover(m, ni, U)
    assert np.array_equal(stress, ss) and np.array_equal(strain, es)     # R4 for the same U: bit for bit
    assert np.abs(strain - strain[:, :1, :]).max() == 0.0                 # G1: every node gets the one Gauss value


def test_example1_analogue_alglib_defaults(solver, oracle):
    """BASELINE config 0 analogue (the Example 1 mesh is missing from the checkout): 15x15x51 G2,
    3 parts / 2 materials, tolerance 1e-6, ALGLIB defaults including the energy-functional stop.  Round 1
    measured 2.8e-9 / 9.6e-8 through the CLI (profiles/r01_config_example1_like.json)."""
    m = mesh.beam(15, 15, 51, n_parts=3, tolerance=1e-6, max_iter=0)
    r = solver.SolverLinearStatics(m)
    o = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-6, maxits=0, merit_check=1))
    assert np.array_equal(r.node_index, o.node_index)
    assert r.cg.terminationtype in (1, 7) and o.stats.cg.terminationtype in (1, 7)      # both "NORMAL" (SolverFunctions.cs:323)
    assert _rel(r.U_full, o.U_full) < 1e-8
    assert np.abs(r.stress - o.stress).max() <= 5e-7 * np.abs(o.stress).max()
    # strict twin at the north_star bars
    m.tolerance = 1e-9
    m.max_iter = 5000
    r = solver.SolverLinearStatics(m, merit_check=0)
    o = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-9, maxits=5000, merit_check=0))
    assert r.cg.terminationtype == 1 and o.stats.cg.terminationtype == 1
    assert _rel(r.U_full, o.U_full) < 1e-10
    assert np.abs(r.stress - o.stress).max() <= 1e-8 * np.abs(o.stress).max()
