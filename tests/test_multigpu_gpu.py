"""Needs two GPUs on the box (skipped otherwise): runs tests/multigpu_worker.py under torchrun with one rank per GPU."""
import os
import subprocess
import sys
import pytest

pytestmark = pytest.mark.gpu


def test_sharded_frames_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(here, "multigpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
