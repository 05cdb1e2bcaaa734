"""Post-processing scalars (Part.Load_Scalar, Part.cs:231-528): numpy oracle known answers on CPU,
GPU kernel against the oracle."""
import numpy as np
import pytest

from oracle import postprocess as PP
from stan_b200 import mesh


def test_oracle_known_answers():
    # uniaxial stress s: principal (s, 0, 0), von Mises = |s|; pure shear t: principal (t, 0, -t), von Mises = sqrt(3) t
    d = np.zeros((2, 3)); d[0] = [3.0, 4.0, 12.0]
    sig = np.array([[100.0, 0, 0, 0, 0, 0], [0, 0, 0, 50.0, 0, 0]])
    eps = np.array([[1e-3, -3e-4, -3e-4, 0, 0, 0], [0, 0, 0, 2e-3, 0, 0]])
    v = PP.node_scalars(d, sig, eps)
    assert v[0, 3] == 13.0                                        # total displacement
    np.testing.assert_allclose(v[0, 10:14], [100.0, 0.0, 0.0, 100.0], atol=1e-12)
    np.testing.assert_allclose(v[1, 10:14], [50.0, 0.0, -50.0, np.sqrt(3) * 50.0], rtol=1e-14, atol=1e-12)
    np.testing.assert_allclose(v[0, 20:23], [1e-3, -3e-4, -3e-4], atol=1e-18)
    np.testing.assert_allclose(v[0, 23], (2.0 / 3.0) * 1.3e-3, rtol=1e-13)   # effective strain
    np.testing.assert_allclose(v[1, 20:23], [2e-3, 0.0, -2e-3], atol=1e-18)  # engineering shear used as is (Part.cs:360)
    m = mesh.beam(2, 2, 2)
    ni = np.arange(m.n_nodes, dtype=np.int32)
    U = np.zeros(m.n_dof); U[0::3] = 1.0
    s = np.zeros((m.n_elem, 8, 6)); s[..., 0] = np.arange(m.n_elem)[:, None] + 1.0
    cell, point = PP.load_scalar(m, ni, U, np.zeros_like(s), s)
    assert cell.dtype == np.float32 and cell.shape == (8, 24, 3) and point.shape == (27, 24)
    assert np.all(cell[:, 4, 0] == np.arange(1, 9)) and np.all(cell[:, 0, :] == 1.0)
    assert point[13, 4] == np.float32(4.5) and point[0, 4] == 1.0  # centre node averages all 8 elements


@pytest.mark.gpu
def test_load_scalar_matches_oracle(oracle):
    from stan_b200.solver import Solver
    m = mesh.beam(5, 4, 12, jitter=True, n_parts=2, tolerance=1e-9)
    with Solver() as s:
        r = s.SolverLinearStatics(m, merit_check=0)
        cell, point, ms = s.Load_Scalar()
    ocell, opoint = PP.load_scalar(m, r.node_index, r.U_full, r.strain, r.stress)
    scale_c = np.abs(ocell).max(axis=(0, 2), keepdims=True) + 1e-30
    scale_p = np.abs(opoint).max(axis=0, keepdims=True) + 1e-30
    assert cell.shape == ocell.shape == (m.n_elem, 24, 3) and point.shape == opoint.shape == (m.n_nodes, 24)
    assert cell.dtype == np.float32 and point.dtype == np.float32
    assert (np.abs(cell - ocell) / scale_c).max() < 2e-6           # float32 storage
    assert (np.abs(point - opoint) / scale_p).max() < 2e-6
    assert ms > 0
