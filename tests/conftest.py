"""pytest plumbing: the ``gpu`` marker, import paths (repo root for ``mlx_video_b200``; ``oracle/`` — the
CPU checker, test infrastructure only), and shared helpers."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(GOLDEN / f"{name}.npz"))
    return load
