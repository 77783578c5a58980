import json
import lzma
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_kitti(i):
    """tests/golden/cloud{i}.xyz.f32.xz -> (n,3) float32 (the reference's test/cloud{i}.bin, xyz only)."""
    with open(os.path.join(GOLDEN, f"cloud{i}.xyz.f32.xz"), "rb") as f:
        return np.frombuffer(lzma.decompress(f.read()), np.float32).reshape(-1, 3).copy()


@pytest.fixture(scope="session")
def kitti():
    return [load_kitti(i) for i in range(1, 5)]


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def pose_delta(A, B):
    """(translation distance [m], rotation angle [rad]) between two 4x4 poses."""
    D = np.linalg.inv(np.asarray(A, np.float64)) @ np.asarray(B, np.float64)
    c = (np.trace(D[:3, :3]) - 1.0) / 2.0
    return float(np.linalg.norm(D[:3, 3])), float(np.arccos(np.clip(c, -1.0, 1.0)))
