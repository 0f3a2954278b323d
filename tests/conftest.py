"""Shared fixtures.  The reference's own fixtures (tests/conftest.py:5-21 there)
are restated as ``toy_cycle_adj`` / ``toy_cycle_csr`` so the drop-in tests read
like the reference's tests."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture
def toy_cycle_adj() -> np.ndarray:
    adj = np.zeros((4, 4))
    for u, v in [(0, 1), (1, 2), (2, 3), (3, 0)]:
        adj[u, v] = adj[v, u] = 1.0
    return adj


@pytest.fixture
def toy_cycle_csr(toy_cycle_adj):
    import scipy.sparse as sp

    return sp.csr_matrix(toy_cycle_adj)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_csr(z, prefix, shape=None):
    import scipy.sparse as sp

    indptr = z[prefix + "_indptr"]
    n = len(indptr) - 1
    return sp.csr_matrix((z[prefix + "_data"], z[prefix + "_indices"].astype(np.int32), indptr.astype(np.int32)),
                         shape=shape or (n, n))
