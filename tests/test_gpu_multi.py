"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): both partitioned modes against the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,port", [("owner", 29541), ("exchange", 29542), ("owner_rows", 29543)])
def test_distributed_assembly_two_gpus(mode, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    env = dict(os.environ, LFGPU_DIST_MODE=mode)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400, env=env)
    assert "DIST_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_distributed_ownership_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29544", os.path.join(ROOT, "tests", "dist_owned_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert "DIST_OWNED_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
