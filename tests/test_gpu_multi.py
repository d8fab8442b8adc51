"""GPU (>= 2 devices): column-sharded operators and drivers against the single-process oracle.
Skipped on a single-GPU box; the same script runs by hand under `gpurun --gpus N` (profiles/multigpu_r01.txt)."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_column_sharded_parity_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTIGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
