import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def checksum(d):
    """Same as tests/golden/make_golden.py:checksum -- guards the seeded generators against drift."""
    tot = 0.0
    for k in sorted(d):
        v = d[k]
        if isinstance(v, np.ndarray):
            a = v.astype(np.float64).ravel()
            tot += float(np.abs(a).sum()) + 1e-3 * float((a * np.arange(1, a.size + 1) % 7).sum())
    return tot


def rel_err(a, b):
    """max|a-b| / max|b| -- the per-tensor metric the 1e-3 relation tolerance is stated in."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
