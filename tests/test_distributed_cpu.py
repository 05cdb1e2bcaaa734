"""N > 1 host-side logic on CPU: world_size 2 over gloo (no GPU)."""
import json
import os
import subprocess
import sys

import numpy as np

from stan_b200 import partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_bounds_cover_all_rows():
    for n, w in [(10, 1), (10, 3), (110691, 8), (10211201, 8), (7, 7)]:
        b = partition.bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1 and np.all(np.diff(b) >= 0)
        assert np.diff(b).max() - np.diff(b).min() <= 1
        q = np.arange(n)
        own = partition.owner(q, b)
        assert np.all((q >= b[own]) & (q < b[own + 1]))


def test_weighted_bounds_balance_blocks():
    from stan_b200 import mesh
    m = mesh.beam(6, 5, 40)
    ni = np.arange(m.n_nodes, dtype=np.int32)                     # x-fastest order: slabs along z like the BFS order
    pairs = partition.node_adjacency(m.conn, ni)
    blocks = np.bincount(pairs[:, 0], minlength=m.n_nodes)        # stored blocks per row
    for w in (2, 3, 8):
        b = partition.weighted_bounds(m.conn, ni, w)
        assert b[0] == 0 and b[-1] == m.n_nodes and np.all(np.diff(b) > 0)
        per_rank = np.add.reduceat(blocks, b[:-1])
        assert per_rank.max() <= 1.05 * per_rank.mean()           # equal node counts would leave the end ranks 3-4 % short
    assert np.array_equal(partition.weighted_bounds(m.conn, ni, 1), [0, m.n_nodes])


def test_gloo_world2_partition_halo_and_reductions():
    r = _torchrun([os.path.join("tests", "dist_worker.py")], 29621)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DIST_OK" in r.stdout


def test_reference_arm_under_torchrun_rank0_only():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--workload",
                   "beam_100k_g2"], 29622)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # rank 1 exits 0 without printing
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["unit"] == "elements/s"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
