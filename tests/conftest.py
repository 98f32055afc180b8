import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One device context per test session.  No fallback: fails if the CUDA library or GPU is missing."""
    import vo_b200
    c = vo_b200.Context(0)
    yield c
    c.close()


from vo_b200.synth import sift_like_descriptors  # noqa: E402,F401  (SURVEY 8d config 4 generator, shared with bench.py)


def correlated_pair(n1, n2, seed, frac=0.5, noise=6.0):
    """Two SIFT-like descriptor sets where a fraction of f1 rows are noisy copies of f2 rows."""
    rng = np.random.default_rng(seed)
    f2 = sift_like_descriptors(n2, seed + 1)
    f1 = sift_like_descriptors(n1, seed + 2)
    k = int(frac * min(n1, n2))
    src = rng.permutation(n2)[:k]
    dst = rng.permutation(n1)[:k]
    f1[dst] = np.clip(np.rint(f2[src] + rng.normal(0, noise, (k, f2.shape[1]))), 0, 255)
    return f1.astype(np.float32), f2
