import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def make_blobs(N, D, K, seed=0, spread=6.0, diag=False):
    """Synthetic mixture in the style of SURVEY.md 8(d): means U(-spread,spread)^D,
    covariances A A^T / D + 0.5 I (or diagonal U(0.5,2)), Dirichlet(5) weights."""
    rng = np.random.default_rng(seed)
    mu = rng.uniform(-spread, spread, size=(K, D))
    w = rng.dirichlet(5.0 * np.ones(K))
    z = rng.choice(K, size=N, p=w)
    X = np.empty((N, D))
    for k in range(K):
        idx = np.nonzero(z == k)[0]
        if diag:
            s = np.sqrt(rng.uniform(0.5, 2.0, size=D))
            X[idx] = mu[k] + rng.normal(size=(idx.size, D)) * s
        else:
            A = rng.normal(size=(D, D))
            C = A @ A.T / D + 0.5 * np.eye(D)
            L = np.linalg.cholesky(C)
            X[idx] = mu[k] + rng.normal(size=(idx.size, D)) @ L.T
    return X, z


def soft_labels(z, K, seed=0, noise=0.3):
    rng = np.random.default_rng(seed)
    q = np.full((z.size, K), noise / K)
    q[np.arange(z.size), z] += 1.0 - noise
    q *= rng.uniform(0.8, 1.2, size=q.shape)
    return q / q.sum(1, keepdims=True)


@pytest.fixture(scope="session")
def testdata():
    d = np.load(os.path.join(ROOT, "tests", "golden", "testdata.npz"))
    return d["X"], d["O"]


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % name))
