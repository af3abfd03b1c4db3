"""Multi-GPU (needs >= 2 GPUs; skipped otherwise): the averaged gradients every rank ends up with — through torch DDP
around the libccx-backed modules, and through CapturedTrainStep's own NCCL buckets inside the CUDA graph — against the
mean over ranks of the ORACLE's per-rank gradients (tests/workers/ddp_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("kind", ["lstm", "transformer", "captured"])
def test_multi_gpu_gradients_equal_oracle_mean_gradients(kind):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "workers", "ddp_worker.py"), kind]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "DDP_OK" in out.stdout
