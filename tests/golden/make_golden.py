#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (``/root/reference`` is not on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules from ``/root/reference`` (with a
three-line stub for the absent ``linear_operator`` package, needed only to get
past ``efficient_graph_gp_sparse/utils_sparse/__init__.py:2``), runs its
samplers and both ``fast_general_grf_kernel``s on small seeded graphs, and
records (a) the outputs and (b) the PCG64 draws each worker consumed, by
calling the reference's ``_init_worker`` + ``_worker_walks`` in-process with
``numpy.random.default_rng`` wrapped in a recording proxy.  The
pooled result and the in-process result are asserted identical before
anything is written.

Nothing in here is product code; the fixtures pin ``oracle/`` to the reference.
"""

import os
import sys
import tempfile
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _install_reference():
    stub_root = tempfile.mkdtemp(prefix="lo_stub_")
    pkg = os.path.join(stub_root, "linear_operator", "operators")
    os.makedirs(pkg)
    open(os.path.join(stub_root, "linear_operator", "__init__.py"), "w").close()
    with open(os.path.join(pkg, "__init__.py"), "w") as fh:
        fh.write("class LinearOperator:\n    def __init__(self, *a, **k):\n        pass\n")
    sys.path.insert(0, stub_root)
    sys.path.insert(0, REF)


class _RecordingRNG:
    """Wraps a numpy Generator; logs every draw the walk loop makes."""

    def __init__(self, rng, log):
        self._rng = rng
        self._log = log

    def random(self):
        u = self._rng.random()
        self._log.append((0, float(u)))
        return u

    def integers(self, n):
        k = self._rng.integers(n)
        self._log.append((1, float(k)))
        return k

    def choice(self, arr):
        # dense sampler: rng.choice(neighbors); log the *position* drawn
        v = self._rng.choice(arr)
        pos = int(np.flatnonzero(np.asarray(arr) == v)[0])
        self._log.append((1, float(pos)))
        return v


def _csr_pack(prefix, m, out):
    m = m.tocsr()
    out[prefix + "_indptr"] = m.indptr.astype(np.int64)
    out[prefix + "_indices"] = m.indices.astype(np.int64)
    out[prefix + "_data"] = m.data.astype(np.float64)
    out[prefix + "_sorted"] = np.array(int(m.has_sorted_indices))


def _record_sparse(module, adj, num_walks, p_halt, L, seed, n_processes):
    """reference workers run in-process with a recording RNG."""
    import numpy.random as npr

    a = adj.tocsr()
    n = a.shape[0]
    module._init_worker(a.indptr, a.indices, a.data.astype(float, copy=False), n)
    real = npr.default_rng
    logs, results = [], []
    try:
        for i, chunk in enumerate(np.array_split(np.arange(n), n_processes)):
            log = []
            npr.default_rng = lambda s, _log=log: _RecordingRNG(real(s), _log)
            results.append(module._worker_walks((chunk.tolist(), num_walks, p_halt, L, (seed or 42) + i, False)))
            logs.append(np.array(log, dtype=np.float64).reshape(-1, 2))
    finally:
        npr.default_rng = real
    return logs, results


def _merge(results, L):
    from collections import defaultdict

    accs = [defaultdict(float) for _ in range(L)]
    for res in results:
        for s in range(L):
            for k, v in res[s].items():
                accs[s][k] += v
    return accs


def _grid(nx, ny):
    def path(n):
        return sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr")

    return (sp.kron(sp.eye(ny), path(nx)) + sp.kron(path(ny), sp.eye(nx))).tocsr()


def _random_graph(n, m, seed, weighted):
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    seen = set()
    while len(seen) < m:
        i, j = int(rng.integers(n)), int(rng.integers(n))
        if i == j or (min(i, j), max(i, j)) in seen:
            continue
        seen.add((min(i, j), max(i, j)))
        w = float(rng.uniform(0.5, 2.0)) if weighted else 1.0
        rows += [i, j]
        cols += [j, i]
        vals += [w, w]
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


def main():
    _install_reference()
    from efficient_graph_gp_sparse.random_walk_samplers_sparse import sparse_sampler as ss
    from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian as lap_sparse
    from efficient_graph_gp_sparse.graph_kernels_sparse.fast_grf_kernel_general import (
        fast_general_grf_kernel as kernel_sparse,
    )
    from efficient_graph_gp.random_walk_samplers import sampler as ds
    from efficient_graph_gp.graph_kernels.utils import get_normalized_laplacian as lap_dense
    from efficient_graph_gp.graph_kernels.fast_grf_kernel_general import fast_general_grf_kernel as kernel_dense

    cycle = np.zeros((4, 4))
    for u, v in [(0, 1), (1, 2), (2, 3), (3, 0)]:
        cycle[u, v] = cycle[v, u] = 1.0

    # ---------------- sparse sampler cases -------------------------------
    # (name, walk graph, W, p, L, seed, n_processes)
    g_iso = _random_graph(40, 45, 7, weighted=True)       # has isolated nodes
    sparse_cases = [
        ("cycle4_raw_p1", sp.csr_matrix(cycle), 5, 0.2, 3, 0, 1),       # tests/test_grf_sparse.py:9-16
        ("cycle4_lap_p2", lap_sparse(sp.csr_matrix(cycle)), 10, 0.2, 3, None, 2),
        ("grid5x5_lap_p3", lap_sparse(_grid(5, 5)), 20, 0.1, 4, 42, 3),
        ("ring32_lap_p8", lap_sparse(sp.csr_matrix(np.roll(np.eye(32), 1, 1) + np.roll(np.eye(32), -1, 1))), 7, 0.1, 5, 3, 8),
        ("gnm40_weighted_iso_lap_p4", lap_sparse(g_iso), 10, 0.1, 4, 11, 4),
        ("gnm40_weighted_raw_p2", g_iso, 6, 0.3, 6, 5, 2),
    ]
    for name, graph, W, p, L, seed, nproc in sparse_cases:
        graph = graph.tocsr()
        pooled = ss.SparseRandomWalk(graph, seed=seed).get_random_walk_matrices(W, p, L, n_processes=nproc)
        logs, results = _record_sparse(ss, graph, W, p, L, seed, nproc)
        accs = _merge(results, L)
        out = {"W": W, "p_halt": p, "L": L, "seed": -1 if seed is None else seed, "n_processes": nproc}
        _csr_pack("graph", graph, out)
        for s in range(L):
            keys = list(accs[s].keys())
            m = sp.csr_matrix(
                (np.array([accs[s][k] for k in keys], dtype=float),
                 (np.array([k[0] for k in keys], dtype=np.int32), np.array([k[1] for k in keys], dtype=np.int32))),
                shape=graph.shape) / W
            assert (m != pooled[s]).nnz == 0 and np.array_equal(m.data, pooled[s].data), name
            assert np.array_equal(m.indices, pooled[s].indices) and np.array_equal(m.indptr, pooled[s].indptr)
            _csr_pack(f"step{s}", pooled[s], out)
        for i, log in enumerate(logs):
            out[f"draws{i}"] = log
        np.savez_compressed(os.path.join(HERE, f"sparse_{name}.npz"), **out)
        print("sparse", name, [m.nnz for m in pooled])

    # ---------------- dense sampler cases --------------------------------
    real = np.random.default_rng
    dense_cases = [
        ("cycle4_raw_seq", cycle, 5, 0.2, 3, 0, 1, False),              # tests/test_grf_dense.py:7-13 (sequential path)
        ("cycle4_raw_seq_ablation", cycle, 5, 0.2, 3, 9, 1, True),
        ("grid5x5_lap_p3", lap_dense(_grid(5, 5).toarray()), 12, 0.1, 4, 42, 3, False),
        ("gnm40_iso_lap_p4", lap_dense(g_iso.toarray()), 8, 0.1, 3, None, 4, False),
    ]
    for name, graph, W, p, L, seed, nproc, ablation in dense_cases:
        n = graph.shape[0]
        pooled = ds.RandomWalk(ds.Graph(graph), seed=seed).get_random_walk_matrices(
            W, p, L, n_processes=nproc, ablation=ablation)
        out = {"W": W, "p_halt": p, "L": L, "seed": -1 if seed is None else seed, "n_processes": nproc,
               "ablation": int(ablation), "graph": graph, "tensor": pooled}
        logs = []
        if nproc == 1 or n < 2 * nproc:
            log = []
            np.random.default_rng = lambda s, _log=log: _RecordingRNG(real(s), _log)
            try:
                again = ds.RandomWalk(ds.Graph(graph), seed=seed).get_random_walk_matrices(
                    W, p, L, n_processes=nproc, ablation=ablation)
            finally:
                np.random.default_rng = real
            assert np.array_equal(again, pooled), name
            logs.append(np.array(log, dtype=np.float64).reshape(-1, 2))
        else:
            ds._init_worker(graph, n)
            results = []
            try:
                for i, chunk in enumerate(np.array_split(np.arange(n), nproc)):
                    log = []
                    np.random.default_rng = lambda s, _log=log: _RecordingRNG(real(s), _log)
                    results.append(ds._worker_walks((chunk.tolist(), W, p, L, (seed or 42) + i, False)))
                    logs.append(np.array(log, dtype=np.float64).reshape(-1, 2))
            finally:
                np.random.default_rng = real
            accs = _merge(results, L)
            again = np.zeros((n, n, L))
            for s in range(L):
                for (i, j), v in accs[s].items():
                    again[i, j, s] = v / W
            assert np.array_equal(again, pooled), name
        for i, log in enumerate(logs):
            out[f"draws{i}"] = log
        np.savez_compressed(os.path.join(HERE, f"dense_{name}.npz"), **out)
        print("dense", name, pooled.shape, float(np.abs(pooled).sum()))

    # ---------------- Laplacians + both fast_general_grf_kernel ----------
    nproc_here = os.cpu_count()
    kern = {"n_processes": nproc_here}
    for name, adj in [("cycle4", sp.csr_matrix(cycle)), ("grid6x4", _grid(6, 4)), ("gnm40w", g_iso)]:
        f = np.array([1.0, 0.5, 0.25])
        kern[name + "_adj"] = adj.toarray()
        _csr_pack(name + "_lap_sparse", lap_sparse(adj), kern)
        kern[name + "_lap_dense"] = lap_dense(adj.toarray())
        kern[name + "_f"] = f
        kern[name + "_K_sparse"] = kernel_sparse(adj, f, walks_per_node=10, p_halt=0.2, max_walk_length=3).toarray()
        kern[name + "_K_dense"] = kernel_dense(adj.toarray(), f, walks_per_node=10, p_halt=0.2, max_walk_length=3)
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **kern)
    print("kernels written (n_processes = os.cpu_count() =", nproc_here, ")")


if __name__ == "__main__":
    main()
