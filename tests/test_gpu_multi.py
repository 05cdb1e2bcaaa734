"""Partitioned solve on 2 GPUs against the single-GPU solve: the default peer-memory data plane (push / wait
kernels storing straight into the neighbours' vector tails, flag-carrying reductions), the variant with the halo
exchange fused into the product kernel, the warp-per-row fallback SpMV, and the NCCL data plane (halo send/recv +
all-reduced dots)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{}, {"STAN_FUSED_HALO": "1"}, {"STAN_SPMV": "0"}, {"STAN_COMM": "nccl"}],
                         ids=["push_wait_kernels", "fused_halo", "fallback_spmv", "nccl"])
def test_two_gpu_partition_matches_single_gpu(env):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join("tools", "multi_gpu_check.py"), "10", "8", "40"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env={**os.environ, **env})
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
