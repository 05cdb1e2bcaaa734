"""CPU oracle against analytic known answers (tests/golden/known_answers.json) and scipy.

The reference ships no tests and cannot run here: parity is unpinned by the
reference, so the oracle is pinned by closed forms, hand-worked cases and an independent direct solve.
"""
import json
import os
from collections import deque

import numpy as np
import pytest

from stan_b200 import mesh

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))
CUBE = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)


def test_extrapolation_matrix(oracle):
    ngp, w, dN, N = oracle.hex8_tables(2)            # FE_Library.cs:91-131
    assert ngp == 8 and w == 1.0
    np.testing.assert_allclose(N[0], GOLD["hex8_g2_extrapolation_row0"], rtol=1e-14)
    np.testing.assert_allclose(N.sum(axis=1), 1.0, rtol=1e-14)
    np.testing.assert_allclose(dN.sum(axis=2), 0.0, atol=1e-16)   # partition of unity
    ngp1, w1, dN1, N1 = oracle.hex8_tables(1)        # FE_Library.cs:63-89
    assert ngp1 == 1 and w1 == 8.0 and np.all(N1 == 1.0)
    assert np.all(np.abs(dN1) == 0.125)


def test_material(oracle):
    D = oracle.elastic_D(210000.0, 0.3)              # Material.cs:31-56
    lam, G = GOLD["unit_cube_steel_lambda"], GOLD["unit_cube_steel_G"]
    np.testing.assert_allclose(D[0, 0], lam + 2 * G, rtol=1e-15)
    np.testing.assert_allclose(D[0, 1], lam, rtol=1e-15)
    np.testing.assert_allclose(np.diag(D)[3:], G, rtol=1e-15)
    assert np.count_nonzero(D) == 12


@pytest.mark.parametrize("etype,nulls", [(2, 6), (1, 18)])
def test_unit_cube_ke(oracle, etype, nulls):
    D = oracle.elastic_D(210000.0, 0.3)
    K = oracle.k_initial(etype, CUBE, D)             # Element.cs:118-155
    assert np.abs(K - K.T).max() < 1e-10 * np.abs(K).max()
    ev = np.linalg.eigvalsh((K + K.T) / 2)
    assert (np.abs(ev) < 1e-9 * ev.max()).sum() == nulls
    if etype == 2:
        np.testing.assert_allclose(K[0, 0], GOLD["unit_cube_steel_K00"], rtol=1e-13)
        assert np.count_nonzero(K) == 576            # no exact zeros
    # rigid translations and rotations are in the null space
    for d in range(3):
        t = np.zeros(24); t[d::3] = 1.0
        assert np.abs(K @ t).max() < 1e-9 * np.abs(K).max()
    rot = np.cross(np.array([0.0, 0.0, 1.0]), CUBE).reshape(24)
    assert np.abs(K @ rot).max() < 1e-9 * np.abs(K).max()


def test_singular_jacobian_is_an_error(oracle):
    flat = CUBE.copy(); flat[:, 2] = 0.0             # MatrixST.cs:298,317 throws on det == 0
    with pytest.raises(ArithmeticError):
        oracle.k_initial(2, flat, oracle.elastic_D(1.0, 0.3))


def _bfs_standard(m):
    n2e = [[] for _ in range(m.n_nodes)]
    for e, c in enumerate(m.conn):
        for n in c:
            if not n2e[n] or n2e[n][-1] != e:
                n2e[n].append(e)
    first = next(n for i in range(1, 7) for n in range(m.n_nodes) if len(n2e[n]) == i)
    idx = -np.ones(m.n_nodes, int); idx[first] = 0
    c, q = 1, deque([first])
    while q:
        v = q.popleft()
        for e in n2e[v]:
            for w in m.conn[e]:
                if idx[w] < 0:
                    idx[w] = c; c += 1; q.append(w)
    return idx


def test_assign_dof(oracle):
    m = mesh.beam(2, 2, 2)
    ni = oracle.assign_dof(m)                        # Database.cs:140-234
    for node, index in GOLD["bfs_2x2x2_first8"].items():
        assert ni[int(node)] == index
    assert sorted(ni) == list(range(m.n_nodes))
    for dims in [(3, 4, 9), (1, 1, 5), (5, 2, 3)]:
        m = mesh.beam(*dims)
        assert np.array_equal(oracle.assign_dof(m), _bfs_standard(m))
    # shuffled element order changes the numbering deterministically and stays a permutation
    rng = np.random.default_rng(0)
    m = mesh.beam(3, 3, 4); m.conn = np.ascontiguousarray(m.conn[rng.permutation(m.n_elem)])
    assert np.array_equal(oracle.assign_dof(m), _bfs_standard(m))


def test_assign_dof_disconnected_and_no_start(oracle):
    m = mesh.beam(1, 1, 1)
    m2 = mesh.beam(1, 1, 1)
    m.xyz = np.vstack([m.xyz, m2.xyz + 5.0]); m.conn = np.vstack([m.conn, m2.conn + 8]).astype(np.int32)
    m.elem_type = np.repeat(m.elem_type, 2); m.elem_mat = np.repeat(m.elem_mat, 2)
    with pytest.raises(RuntimeError):                # reference: index out of range (Database.cs:218)
        oracle.assign_dof(m)


def test_spc_reduction_and_rhs(oracle):
    m = mesh.beam(1, 1, 2)
    ni = np.arange(m.n_nodes, dtype=np.int32)[::-1].copy()   # any numbering works for this unit
    m.spc_node = np.array([0, 2, 2], np.int32)
    m.spc_val = np.array([[1, 0, 1], [0, 1, 0], [0, 1, 0]], float)   # duplicates collapse (Distinct)
    red, nfix = oracle.spc_reduction(m, ni)          # Solver.cs:104-132
    assert nfix == 3
    fixed = sorted([3 * ni[0] + 0, 3 * ni[0] + 2, 3 * ni[2] + 1])
    exp = np.zeros(m.n_dof, np.int32); c = 0
    for i in range(m.n_dof):
        if i in fixed: exp[i] = -1; c += 1
        else: exp[i] = c
    assert np.array_equal(red, exp)
    m.load_node = np.array([0, 5, 5], np.int32)
    m.load_val = np.array([[7, 8, 9], [1, 2, 3], [10, 20, 30]], float)
    F = oracle.build_rhs(m, ni, red)                 # Solver.cs:136-152 (fixed DOFs skipped, += accumulates)
    full = oracle.include_bc_dof(red, F)             # SolverFunctions.cs:520-538
    assert full[3 * ni[0] + 0] == 0 and full[3 * ni[0] + 1] == 8 and full[3 * ni[0] + 2] == 0
    assert list(full[3 * ni[5]: 3 * ni[5] + 3]) == [11, 22, 33]
    assert full.sum() == 8 + 66


def test_assembly_pattern_and_values(oracle):
    m = mesh.beam(3, 3, 3, jitter=True)
    ni = oracle.assign_dof(m)
    red, nfix = oracle.spc_reduction(m, ni)
    K = oracle.assemble_upper(m, ni, red)            # SolverFunctions.cs:117-180, :275
    Kp = oracle.assemble_upper(m, ni, red, prune=True)
    rp, col, val = K.arrays()
    assert K.exact_zero == 0 and Kp.nnz == K.nnz     # jitter: structural == ALGLIB-pruned (Appendix C)
    assert K.n == m.n_dof - nfix
    for r in range(K.n):
        c = col[rp[r]:rp[r + 1]]
        assert c[0] == r and np.all(np.diff(c) > 0)  # upper incl. diagonal, ascending
    # same matrix from an independent dense assembly in full space
    A = np.zeros((m.n_dof, m.n_dof))
    for e in range(m.n_elem):
        D = oracle.elastic_D(m.mat_E[m.elem_mat[e]], m.mat_nu[m.elem_mat[e]])
        Ke = oracle.k_initial(int(m.elem_type[e]), m.xyz[m.conn[e]], D)
        dofs = (3 * ni[m.conn[e]][:, None] + np.arange(3)[None, :]).ravel()
        A[np.ix_(dofs, dofs)] += Ke
    free = np.where(red != -1)[0]
    Ar = A[np.ix_(free, free)]
    np.testing.assert_allclose(K.to_scipy_full().toarray(), (np.triu(Ar) + np.triu(Ar, 1).T), rtol=1e-12, atol=1e-9)
    # regular grid: ALGLIB-style pruning removes exact zeros, so the stored pattern shrinks
    mr = mesh.beam(3, 3, 3)
    nir = oracle.assign_dof(mr); redr, _ = oracle.spc_reduction(mr, nir)
    Ks, Kq = oracle.assemble_upper(mr, nir, redr), oracle.assemble_upper(mr, nir, redr, prune=True)
    assert Ks.exact_zero > 0 and Kq.nnz == Ks.nnz - Ks.exact_zero


def test_sym_spmv_matches_scipy(oracle):
    m = mesh.beam(2, 3, 4, jitter=True)
    ni = oracle.assign_dof(m); red, _ = oracle.spc_reduction(m, ni)
    K = oracle.assemble_upper(m, ni, red)
    x = np.random.default_rng(1).standard_normal(K.n)
    np.testing.assert_allclose(oracle.sym_spmv(K, x), K.to_scipy_full() @ x, rtol=1e-12, atol=1e-8)


def _solve(oracle, m, **kw):
    ni = oracle.assign_dof(m); red, _ = oracle.spc_reduction(m, ni)
    F = oracle.build_rhs(m, ni, red)
    K = oracle.assemble_upper(m, ni, red)
    x, rep = oracle.lincg(K, F, oracle.cg_opts(**kw))
    return ni, red, F, K, x, rep


def test_lincg_against_direct_solve(oracle):
    import scipy.sparse.linalg as spl
    m = mesh.beam(4, 4, 50)
    ni, red, F, K, x, rep = _solve(oracle, m, epsf=1e-8)
    xs = spl.spsolve(K.to_scipy_full().tocsc(), F)
    assert rep.terminationtype == 1                  # ||r|| <= EpsF ||b||
    assert np.sqrt(rep.r2) <= 1e-8 * rep.bnorm
    assert rep.nmv == 1 + rep.iterationscount + rep.iterationscount // 10
    assert 150 < rep.iterationscount < 260           # ~4.1 x nz (202 measured there)
    assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-10
    # cantilever tip deflection vs Euler-Bernoulli + shear, within discretisation error
    U = oracle.include_bc_dof(red, x)
    tip = U[3 * ni[m.load_node]].mean()
    EI, L, G, A = 210000.0 * 4**4 / 12, 50.0, 80769.23, 16.0
    beam_theory = 1000 * L**3 / (3 * EI) + 1000 * L / (5.0 / 6.0 * G * A)
    assert abs(tip - beam_theory) / beam_theory < 0.06


def test_lincg_termination_codes(oracle):
    m = mesh.beam(2, 2, 10)
    *_, rep = _solve(oracle, m, epsf=1e-8, maxits=5)
    assert rep.terminationtype == 5 and rep.iterationscount == 5
    *_, rep = _solve(oracle, m, epsf=1e-30)          # unreachable: energy functional stalls -> 7
    assert rep.terminationtype == 7 and rep.iterationscount % 10 == 0
    *_, rep = _solve(oracle, m, epsf=0.0, maxits=0)  # lincgsetcond: both zero -> EpsF = 1e-6
    assert rep.terminationtype == 1 and np.sqrt(rep.r2) <= 1e-6 * rep.bnorm
    m.load_val[:] = 0.0
    *_, x, rep = _solve(oracle, m, epsf=1e-8)
    assert rep.terminationtype == 1 and rep.iterationscount == 0 and not x.any()
    # G1 with an even section is singular/inconsistent: never reports type 1
    g1 = mesh.beam(2, 2, 6, elem_type=mesh.HEX8_G1)
    *_, rep = _solve(oracle, g1, epsf=1e-8, maxits=500)
    assert rep.terminationtype != 1


def test_cholesky_skyline_against_dense_and_theory(oracle):
    """LinearSolver_Cholesky restatement (SolverFunctions.cs:332-444): pinned by numpy's dense Cholesky,
    the envelope definition, beam theory, and the -3 / zeros contract for a matrix that is not SPD."""
    m = mesh.beam(4, 4, 50)
    ni = oracle.assign_dof(m); red, _ = oracle.spc_reduction(m, ni)
    F = oracle.build_rhs(m, ni, red)
    K = oracle.assemble_upper(m, ni, red)
    x, tt, env = oracle.cholesky_skyline(K, F)
    A = K.to_scipy_full().toarray()
    L = np.linalg.cholesky(A)
    xd = np.linalg.solve(L.T, np.linalg.solve(L, F))
    assert tt == 1 and np.linalg.norm(x - xd) / np.linalg.norm(xd) < 1e-11
    rp, col, _ = K.arrays()
    first = np.arange(K.n)
    for i in range(K.n):
        c = col[rp[i]:rp[i + 1]]
        first[c] = np.minimum(first[c], i)
    assert env == int((np.arange(K.n) - first + 1).sum())         # sparseconverttosks envelope
    U = oracle.include_bc_dof(red, x)
    tip = U[3 * ni[m.load_node]].mean()
    EI, Lb, G, Ar = 210000.0 * 4**4 / 12, 50.0, 80769.23, 16.0
    theory = 1000 * Lb**3 / (3 * EI) + 1000 * Lb / (5.0 / 6.0 * G * Ar)
    assert abs(tip - theory) / theory < 0.06
    xc, rep = oracle.lincg(K, F, oracle.cg_opts(epsf=1e-10, maxits=5000, merit_check=0))
    assert rep.terminationtype == 1 and np.linalg.norm(x - xc) / np.linalg.norm(x) < 1e-8
    m.mat_E = -m.mat_E                                            # negative definite: no factor
    Kn = oracle.assemble_upper(m, ni, red)
    xn, tt, _ = oracle.cholesky_skyline(Kn, F)
    assert tt == -3 and not xn.any()                              # "filled by zeros" (SolverFunctions.cs:420)


def test_recovery_patch_test(oracle):
    m = mesh.beam(3, 3, 3, jitter=True)
    ni = oracle.assign_dof(m)
    G = np.array([[1e-3, 2e-4, -1e-4], [3e-4, -2e-3, 5e-4], [-2e-4, 1e-4, 1.5e-3]])
    u = m.xyz @ G.T
    U = np.zeros(m.n_dof)
    for d in range(3):
        U[3 * ni + d] = u[:, d]
    strain, stress = oracle.recover(m, ni, U)        # Element.cs:211-246, 257-267
    eps = np.array([G[0, 0], G[1, 1], G[2, 2], G[0, 1] + G[1, 0], G[1, 2] + G[2, 1], G[0, 2] + G[2, 0]])
    np.testing.assert_allclose(strain, np.broadcast_to(eps, strain.shape), rtol=1e-11, atol=1e-16)
    np.testing.assert_allclose(stress, np.broadcast_to(oracle.elastic_D(210000.0, 0.3) @ eps, stress.shape), rtol=1e-11)
    m.elem_type[:] = mesh.HEX8_G1                    # G1: intent N[g][i] = 1 (reference throws, §8a R4)
    s1, _ = oracle.recover(m, ni, U)
    np.testing.assert_allclose(s1, np.broadcast_to(eps, s1.shape), rtol=1e-11, atol=1e-16)


def test_whole_path_driver(oracle):
    m = mesh.beam(3, 3, 12, n_parts=2)
    r = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-8))
    assert r.stats.cg.terminationtype in (1, 7) and r.stats.n_free == m.n_dof - 3 * 16
    r2 = oracle.linear_statics(m, oracle.cg_opts(epsf=1e-8, parallel_spmv=1))
    np.testing.assert_allclose(r2.U_full, r.U_full, rtol=1e-7, atol=1e-12)
    assert np.abs(r.U_full[3 * r.node_index[m.spc_node]]).max() == 0.0


def test_mesh_generator_and_bdf(tmp_path):
    m = mesh.beam(2, 3, 4, jitter=True)
    assert m.n_nodes == 3 * 4 * 5 and m.n_elem == 24 and m.conn.max() == m.n_nodes - 1
    r = mesh.beam(2, 3, 4)
    moved = np.abs(m.xyz - r.xyz).max(axis=1) > 0
    assert 0 < moved.sum() <= 1 * 2 * 3 and np.abs(m.xyz - r.xyz).max() <= 0.1
    np.testing.assert_allclose(m.load_val[:, 0].sum(), 1000.0)
    p = tmp_path / "m.bdf"
    mesh.write_bdf(r, str(p))
    lines = p.read_text().splitlines()
    assert sum(l.startswith("GRID") for l in lines) == r.n_nodes
    assert sum(l.startswith("CHEXA") for l in lines) == r.n_elem


def test_alglib_restatements_on_hand_worked_matrices(oracle):
    """Hand-worked known answers for the two ALGLIB restatements, no mesh involved.
    A = [[4,2,0,0],[2,5,3,0],[0,3,10,1],[0,0,1,2]]: U^T U by hand is u11 = 2, u12 = 1, u22 = sqrt(5-1) = 2,
    u23 = 3/2, u33 = sqrt(10 - 9/4), u34 = 1/u33, u44 = sqrt(2 - 1/7.75); skyline envelope 1+2+2+2 = 7.
    CG on a 2x2 SPD system reaches the exact solution in two iterations (Jacobi-preconditioned or not)."""
    A = np.array([[4.0, 2, 0, 0], [2, 5, 3, 0], [0, 3, 10, 1], [0, 0, 1, 2]])
    x_true = np.array([1.0, -1.0, 2.0, 0.5])
    K = oracle.UpperCsr.from_dense_upper(A)
    x, tt, env = oracle.cholesky_skyline(K, A @ x_true)
    assert tt == 1 and env == 7
    np.testing.assert_allclose(x, x_true, rtol=0, atol=4e-16 * 8)
    np.testing.assert_allclose(oracle.sym_spmv(K, x_true), A @ x_true, rtol=0, atol=1e-15)
    # indefinite: the second pivot 1 - 4 < 0 -> -3 and zeros
    xi, tt, _ = oracle.cholesky_skyline(oracle.UpperCsr.from_dense_upper([[1.0, 2.0], [2.0, 1.0]]), np.array([1.0, 1.0]))
    assert tt == -3 and not xi.any()
    B = np.array([[4.0, 1.0], [1.0, 3.0]])
    b = np.array([1.0, 2.0])
    xc, rep = oracle.lincg(oracle.UpperCsr.from_dense_upper(B), b, oracle.cg_opts(epsf=1e-14, maxits=10, merit_check=0))
    np.testing.assert_allclose(xc, [1.0 / 11.0, 7.0 / 11.0], rtol=1e-14)
    assert rep.iterationscount == 2 and rep.terminationtype == 1 and rep.nmv == 3
