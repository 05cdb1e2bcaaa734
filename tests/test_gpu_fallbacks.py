"""The non-default kernel paths stay correct: assembly in row chunks (STAN_ASM_CHUNK_ROWS), the warp-per-row SpMV that
is the automatic fallback for very wide rows (STAN_SPMV=0) and the intermediate bulk-copy variants.
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
