import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN_DIR, f"{tag}.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def golden_tags(kind=None):
    tags = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
    if kind is None:
        return tags
    return [t for t in tags if load_golden(t)[0]["kind"] == kind]


def rel_err(a, b):
    """max|a-b| / max|b| — the parity metric of SURVEY.md §8c."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
