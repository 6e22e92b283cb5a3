"""pytest configuration: the ``gpu`` marker and import paths.

``-m "not gpu"`` runs on the CPU-only build container (oracle vs golden vectors,
host logic, C-ABI symbol checks, gloo sharding); ``-m gpu`` runs the parity
tests proper on a B200 through the C-ABI.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "monocular-visual-slam_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
