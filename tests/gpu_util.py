"""Helpers shared by the -m gpu parity tests."""

import numpy as np
import scipy.sparse as sp


def csr_bits_equal(a, b) -> bool:
    a, b = a.tocsr(), b.tocsr()
    return (a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(np.asarray(a.data, dtype=np.float64).view(np.int64),
                               np.asarray(b.data, dtype=np.float64).view(np.int64)))


def grid_graph(nx, ny):
    def path(n):
        return sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr")

    return (sp.kron(sp.eye(ny), path(nx)) + sp.kron(path(ny), sp.eye(nx))).tocsr()


def ring_graph(n):
    i = np.arange(n)
    return sp.csr_matrix((np.ones(2 * n), (np.r_[i, i], np.r_[(i + 1) % n, (i - 1) % n])), shape=(n, n))


def random_graph(n, m, seed, weighted=False):
    rng = np.random.default_rng(seed)
    i = rng.integers(0, n, size=m)
    j = rng.integers(0, n, size=m)
    keep = i != j
    i, j = i[keep], j[keep]
    w = rng.uniform(0.5, 2.0, size=i.size) if weighted else np.ones(i.size)
    a = sp.coo_matrix((w, (i, j)), shape=(n, n)).tocsr()
    a = a.maximum(a.T).tocsr()
    a.sort_indices()
    return a


def powerlaw_graph(n, m, seed):
    """R-MAT-ish: endpoints drawn with probability ~ 1/rank (hubs), symmetrised."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, n + 1)
    p /= p.sum()
    i = rng.choice(n, size=m, p=p)
    j = rng.integers(0, n, size=m)
    keep = i != j
    a = sp.coo_matrix((np.ones(keep.sum()), (i[keep], j[keep])), shape=(n, n)).tocsr()
    a.data[:] = 1.0
    a = a.maximum(a.T).tocsr()
    a.sort_indices()
    return a
