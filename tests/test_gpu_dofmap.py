"""Database.AssignDOF on the device (dofmap_gpu.cu): the level-synchronous traversal must give the
reference's FIFO numbering bit for bit — against the oracle and against the serial host traversal —
including shuffled element order, degenerate elements, many elements around one node, the narrow-mesh
hand-over and the error cases (Database.cs:140-234)."""
import numpy as np
import pytest

from stan_b200 import mesh, native
from stan_b200.solver import Solver

pytestmark = pytest.mark.gpu


def _shuffled(m, seed=0):
    perm = np.random.default_rng(seed).permutation(m.n_elem)
    m.conn = np.ascontiguousarray(m.conn[perm]); m.elem_type = m.elem_type[perm]
    m.elem_mat = m.elem_mat[perm]; m.elem_pid = m.elem_pid[perm]
    return m


def _numbering(m, mode, monkeypatch):
    monkeypatch.setenv("STAN_DOF", mode)
    with Solver() as s:
        s.SetModel(m)
        return s.AssignDOF()


@pytest.mark.parametrize("make", [
    lambda: mesh.beam(1, 1, 1), lambda: mesh.beam(2, 3, 4), lambda: mesh.beam(12, 9, 30, jitter=True),
    lambda: _shuffled(mesh.beam(10, 11, 12), 5), lambda: mesh.beam(40, 40, 40), lambda: _shuffled(mesh.beam(33, 20, 17), 1),
    lambda: mesh.polar_disk(24, 5, 4), lambda: _shuffled(mesh.polar_disk(12, 3, 6), 2), lambda: mesh.beam(2, 2, 400),
])
def test_device_numbering_is_the_reference_numbering(oracle, monkeypatch, make):
    m = make()
    want = oracle.assign_dof(m)
    gpu = _numbering(m, "gpu", monkeypatch)
    host = _numbering(m, "host", monkeypatch)
    assert np.array_equal(gpu, want) and np.array_equal(host, want)
    assert np.array_equal(np.sort(gpu), np.arange(m.n_nodes))        # a permutation


def test_device_numbering_errors(monkeypatch):
    monkeypatch.setenv("STAN_DOF", "gpu")
    a, b = mesh.beam(3, 3, 3), mesh.beam(3, 3, 3)
    n0 = a.n_nodes
    a.xyz = np.vstack([a.xyz, b.xyz + 50.0]); a.conn = np.vstack([a.conn, b.conn + n0]).astype(np.int32)
    a.elem_type = np.repeat(a.elem_type, 2); a.elem_mat = np.repeat(a.elem_mat, 2)
    with Solver() as s:
        s.SetModel(a)
        with pytest.raises(native.StanError) as ei:
            s.AssignDOF()
        assert ei.value.code == native.E_DOFMAP and "disconnected" in str(ei.value)
        # the handle survives and numbers a good mesh afterwards
        m = mesh.beam(5, 5, 5)
        s.SetModel(m)
        assert np.array_equal(np.sort(s.AssignDOF()), np.arange(m.n_nodes))


def test_large_mesh_takes_the_device_path_by_default(oracle, monkeypatch):
    monkeypatch.delenv("STAN_DOF", raising=False)
    m = mesh.beam(100, 100, 100)                                     # 1.03 M nodes: above the automatic threshold
    with Solver() as s:
        s.SetModel(m)
        l0 = s.kernel_launches()
        ni = s.AssignDOF()
        assert s.kernel_launches() - l0 > 100                        # kernels ran: it was not the host traversal
    assert np.array_equal(ni, oracle.assign_dof(m))
