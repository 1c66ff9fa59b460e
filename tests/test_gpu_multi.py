"""GPU, >= 2 devices: sequence-parallel / CFG-parallel parity (scripts/sp_check.py under torchrun)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_parallel_parity(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), "scripts/sp_check.py"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]


def test_parallel_parity_with_folded_barriers():
    """The opt-in form of the cross-GPU flag barrier (ltxb_peer_sync: run in the prologue of the consumer kernel — the
    attention kernel and the out-projection GEMM — instead of as a launch of its own) gives the same parity."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29520", "scripts/sp_check.py"]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, LTXB_FOLD_BARRIERS="1"))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
