"""Multi-GPU (needs >= 2 GPUs; skipped otherwise): DDP over NCCL wraps the libccx-backed modules unchanged."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["lstm", "transformer"])
def test_ddp_gradients_equal_mean_of_local_gradients(kind):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "workers", "ddp_worker.py"), kind]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "DDP_OK" in out.stdout
