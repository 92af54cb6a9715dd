"""Multi-GPU TSQR on real hardware: 2 ranks over NCCL (skipped on a single-GPU box)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n", [(16384, 256), (65536, 1024)])
def test_tsqr_two_gpus(m, n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           str(ROOT / "tests" / "helpers" / "tsqr_worker.py"), str(m), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
