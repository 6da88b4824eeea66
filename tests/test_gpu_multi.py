"""Multi-GPU paths on real GPUs (skipped on single-GPU boxes): NCCL sharding and the fused peer-memory all-reduce."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_fused_p2p_allreduce_and_sharding():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "p2p_check.py"), "30", "48"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "P2P_CHECK OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
