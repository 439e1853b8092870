"""Launches tests/multigpu_parity.py under torchrun when at least two GPUs are visible."""
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_sharded_training_matches_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "multigpu_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTIGPU_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
