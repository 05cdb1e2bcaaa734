"""The non-default kernel paths stay correct: assembly in row chunks (STAN_ASM_CHUNK_ROWS), the warp-per-row SpMV that
is the automatic fallback for very wide rows (STAN_SPMV=0), the intermediate bulk-copy variants, and the device-side
CG timeline (STAN_CG_TRACE).
Each runs __graft_entry__.smoke() — assemble, CG, recovery checked against the oracle — in a fresh
process because the selection is read once per process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"STAN_ASM_CHUNK_ROWS": "100", "STAN_SPMV": "0"}, {"STAN_SPMV": "1"}, {"STAN_SPMV": "2"},
                                 {"STAN_SPMV": "3"}])
def test_alternative_kernel_paths(env):
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT,
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_cg_trace_dump_and_timeline(tmp_path):
    """STAN_CG_TRACE: the CG kernels stamp %globaltimer at their phase boundaries; the dump is one
    "ns,iteration,code" line per event and tools/cg_timeline.py turns it into the per-transition table."""
    env = dict(os.environ, STAN_CG_TRACE="512", STAN_CG_TRACE_FROM="20", STAN_CG_TRACE_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "smoke ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]      # tracing changes no result
    f = tmp_path / "stan_cg_trace_rank0.csv"
    ev = [tuple(int(v) for v in line.split(",")) for line in f.read_text().splitlines()]
    assert 100 <= len(ev) <= 512
    assert all(k >= 20 for _, k, _ in ev)                                   # nothing before STAN_CG_TRACE_FROM
    assert {1, 2, 3, 4, 5, 6, 7} <= {c for _, _, c in ev}                   # product, update and direction marks
    begins = [t for t, _, c in ev if c == 1]
    assert begins == sorted(begins) and begins[-1] > begins[0]
    t = subprocess.run([sys.executable, "tools/cg_timeline.py", str(tmp_path)], cwd=ROOT, capture_output=True, text=True)
    assert t.returncode == 0 and "| spmv begin | spmv local sum done |" in t.stdout and "Iteration period" in t.stdout
