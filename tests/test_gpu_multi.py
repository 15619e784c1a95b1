"""Multi-GPU parity (-m gpu, needs >= 2 devices; skipped otherwise): launches tests/multi_gpu_worker.py under torchrun with two
ranks and requires every sharded mode — peer-store gather, ncclReduce gather, the pipelined frame ring, sharded views — to equal
the single-GPU frame bit for bit (the result must not depend on the partition)."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("ranks", [2, 4])
def test_sharded_modes_equal_single_gpu(ranks):
    if _device_count() < ranks:
        pytest.skip(f"needs {ranks} CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + ranks), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "MULTI-GPU CHECK PASSED" in p.stdout, p.stdout[-4000:]
