"""Edge cases at the C-ABI boundary: empty inputs, duplicate boundary-condition entries,
an unconstrained (singular) system, and re-use of one handle across different models."""
import numpy as np
import pytest

from stan_b200 import mesh, native
from stan_b200.solver import Solver

pytestmark = pytest.mark.gpu


def test_empty_inputs_are_rejected():
    with Solver() as s:
        m = mesh.beam(1, 1, 1)
        m.conn = m.conn[:0]; m.elem_type = m.elem_type[:0]; m.elem_mat = m.elem_mat[:0]
        with pytest.raises(native.StanError) as ei:
            s.SetModel(m)
        assert ei.value.code == native.E_ARG
        m = mesh.beam(1, 1, 1)
        m.elem_type[:] = 9                                       # TET4/PENTA6 are not on this path
        with pytest.raises(native.StanError) as ei:
            s.SetModel(m)
        assert ei.value.code == native.E_ARG
        m = mesh.beam(1, 1, 1)
        m.elem_mat[:] = 3                                        # MatLib has one entry
        s.SetModel(m); s.AssignDOF()
        with pytest.raises(native.StanError) as ei:
            s.ParallelAssembly_K()
        assert ei.value.code == native.E_ARG


def test_duplicate_bc_entries_accumulate_like_the_reference(oracle):
    m = mesh.beam(3, 3, 6, jitter=True, tolerance=1e-10)
    m.spc_node = np.concatenate([m.spc_node, m.spc_node[:5]]).astype(np.int32)          # Distinct() collapses them
    m.spc_val = np.vstack([m.spc_val, np.ones((5, 3))])
    m.load_node = np.concatenate([m.load_node, m.load_node[:4], m.spc_node[:2]]).astype(np.int32)
    extra = np.array([[1.0, 2.0, 3.0]] * 4 + [[9.0, 9.0, 9.0]] * 2)                     # += on loaded nodes; loads on fixed DOFs dropped
    m.load_val = np.vstack([m.load_val, extra])
    with Solver() as s:
        s.SetModel(m); ni = s.AssignDOF(); s.ParallelAssembly_K()
        red, nfix = oracle.spc_reduction(m, ni)
        assert np.array_equal(s.nDOF_reduction(), red) and nfix == 3 * 16
        assert np.array_equal(s.F(), oracle.build_rhs(m, ni, red))
        rep = s.LinearSolver_CG(merit_check=0, IterMax=3000)
        K = oracle.assemble_upper(m, ni, red)
        xo, _ = oracle.lincg(K, oracle.build_rhs(m, ni, red), oracle.cg_opts(epsf=1e-10, merit_check=0, maxits=3000))
        assert rep.terminationtype == 1
        assert np.linalg.norm(s.Exclude_BC_DOF() - xo) / np.linalg.norm(xo) < 1e-10


def test_unconstrained_model_terminates_with_a_code():
    m = mesh.beam(3, 3, 6, tolerance=1e-8)
    m.spc_node = m.spc_node[:0]; m.spc_val = m.spc_val[:0]       # rigid-body modes: K is singular
    with Solver() as s:
        s.SetModel(m); s.AssignDOF()
        a = s.ParallelAssembly_K()
        assert a.n_fixed == 0
        rep = s.LinearSolver_CG(IterMax=300)
        assert rep.terminationtype in (5, 7, -4, -5) and rep.iterationscount <= 300     # never "converged"


def test_one_handle_many_models(oracle):
    with Solver() as s:
        for dims, et in [((2, 2, 3), mesh.HEX8_G2), ((6, 5, 9), mesh.HEX8_G2), ((3, 3, 4), mesh.HEX8_G1), ((1, 1, 1), mesh.HEX8_G2)]:
            m = mesh.beam(*dims, elem_type=et, jitter=True, tolerance=1e-8)
            m.max_iter = 500
            r = s.SolverLinearStatics(m)
            o = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-8, maxits=500))
            assert np.array_equal(r.node_index, o.node_index)
            if r.cg.terminationtype == 1 and o.stats.cg.terminationtype == 1:
                assert np.linalg.norm(r.U_full - o.U_full) <= 1e-7 * np.linalg.norm(o.U_full)


def test_results_through_every_door(oracle):
    """U by DOF, U per node (Solver.cs:171-178, gathered on the device), the rows a rank owns, and the same results
    through page-locked buffers from stan_host_alloc: all the same numbers."""
    m = mesh.beam(5, 4, 12, jitter=True, n_parts=2, tolerance=1e-9)
    with Solver() as s, Solver(pinned_results=True) as sp:
        r = s.SolverLinearStatics(m, merit_check=0)
        assert np.array_equal(r.disp, r.U_full.reshape(-1, 3)[r.node_index])      # Node.dU_buffer[i] = U_Full[DOF[i]]
        assert np.array_equal(s.displacements_local().ravel(), r.U_full)          # one GPU owns every row
        mp = sp.pinned_model(m)
        rp = sp.SolverLinearStatics(mp, node_index=r.node_index, merit_check=0)
        for a, b in ((rp.U_full, r.U_full), (rp.disp, r.disp), (rp.strain, r.strain), (rp.stress, r.stress)):
            assert np.array_equal(a, b)
        first = rp.stress
        rp2 = sp.SolverLinearStatics(mp, node_index=r.node_index, merit_check=0)
        assert rp2.stress is first                                                # pinned result buffers are reused


def test_model_checks_run_on_the_device():
    with Solver() as s:
        m = mesh.beam(3, 3, 4)
        bad = mesh.beam(3, 3, 4); bad.conn = bad.conn.copy(); bad.conn[17, 5] = -4
        with pytest.raises(native.StanError) as ei:
            s.SetModel(bad)
        assert ei.value.code == native.E_ARG and "element 17" in str(ei.value) and "-4" in str(ei.value)
        bad = mesh.beam(3, 3, 4); bad.elem_mat = bad.elem_mat.copy(); bad.elem_mat[20] = -1
        with pytest.raises(native.StanError) as ei:
            s.SetModel(bad)
        assert ei.value.code == native.E_ARG and "element 20" in str(ei.value)
        s.SetModel(m)
        with pytest.raises(native.StanError) as ei:                               # a rejected mesh leaves no mesh behind
            s.SetModel(bad)
        with pytest.raises(native.StanError) as ei:
            s.AssignDOF()
        assert ei.value.code == native.E_STATE
        s.SetModel(m)
        ni = s.AssignDOF()
        dup = ni.copy(); dup[3] = dup[4]
        with pytest.raises(native.StanError) as ei:
            s.SetDOF(dup)
        assert ei.value.code == native.E_ARG and "permutation" in str(ei.value)
        out = ni.copy(); out[0] = m.n_nodes
        with pytest.raises(native.StanError) as ei:
            s.SetDOF(out)
        assert ei.value.code == native.E_ARG
        with pytest.raises(native.StanError) as ei:                               # and no DOF map either
            s.ParallelAssembly_K()
        assert ei.value.code == native.E_STATE
        s.SetDOF(ni)
        assert s.ParallelAssembly_K().n_dof == m.n_dof
