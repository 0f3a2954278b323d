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

import hashlib
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


def _ring_csr(n, diag_v, off_v, adjacency=False):
    """CSR of the n-ring: adjacency (unit weights) or its normalized Laplacian with the given two values,
    columns sorted -- built directly, no SpGEMM."""
    i = np.arange(n, dtype=np.int64)
    if adjacency:
        cols = np.sort(np.stack([(i - 1) % n, (i + 1) % n], axis=1), axis=1)
        vals = np.ones((n, 2))
    else:
        cols = np.stack([(i - 1) % n, i, (i + 1) % n], axis=1)
        vals = np.stack([np.full(n, off_v), np.full(n, diag_v), np.full(n, off_v)], axis=1)
        order = np.argsort(cols, axis=1, kind="stable")
        cols = np.take_along_axis(cols, order, axis=1)
        vals = np.take_along_axis(vals, order, axis=1)
    k = cols.shape[1]
    return sp.csr_matrix((vals.ravel(), cols.ravel().astype(np.int32), (np.arange(n + 1, dtype=np.int64) * k).astype(np.int32)),
                         shape=(n, n))


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
        # production walk counts: every instantiation of the CUDA walker replays the reference
        # (W <= 32: 1 key per lane, 40: 2, 100: 4 -- BASELINE's setting --, 130: 8, 300: CTA per start node)
        ("ring60_lap_W100_p3", lap_sparse(sp.csr_matrix(np.roll(np.eye(60), 1, 1) + np.roll(np.eye(60), -1, 1))),
         100, 0.1, 5, 42, 3),
        ("gnm50_weighted_lap_W40_p2", lap_sparse(_random_graph(50, 120, 21, weighted=True)), 40, 0.1, 4, 7, 2),
        ("grid6x6_lap_W130_p4", lap_sparse(_grid(6, 6)), 130, 0.15, 3, 13, 4),
        ("gnm30_weighted_lap_W300_p2", lap_sparse(_random_graph(30, 70, 5, weighted=True)), 300, 0.1, 4, 1, 2),
        # W x L beyond one CTA's shared memory (the reference's wind experiment: W = 8192, L = 5; its ablation
        # notebook: W = 10000, L = 10): the walker's fallbacks -- fewer start nodes per CTA (W = 200, L = 30: two
        # warps; W = 250, L = 40: one), one walk length at a time (W = 2100, L = 10; W = 256, L = 80; W = 8192, L = 5)
        ("gnm16_weighted_lap_W200_L30_p2", lap_sparse(_random_graph(16, 40, 31, weighted=True)), 200, 0.05, 30, 3, 2),
        ("ring12_lap_W250_L40_p1", lap_sparse(sp.csr_matrix(np.roll(np.eye(12), 1, 1) + np.roll(np.eye(12), -1, 1))),
         250, 0.04, 40, 8, 1),
        ("gnm12_weighted_lap_W2100_L10_p2", lap_sparse(_random_graph(12, 30, 32, weighted=True)), 2100, 0.1, 10, 4, 2),
        ("grid3x3_lap_W256_L80_p1", lap_sparse(_grid(3, 3)), 256, 0.03, 80, 6, 1),
        ("gnm6_weighted_lap_W8192_L5_p2", lap_sparse(_random_graph(6, 12, 33, weighted=True)), 8192, 0.1, 5, 9, 2),
    ]
    only = os.environ.get("GOLDEN_ONLY")     # regenerate just the sparse cases whose name contains this
    for name, graph, W, p, L, seed, nproc in sparse_cases:
        if only and only not in name:
            continue
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
            if log.shape[0] <= 4000:
                out[f"draws{i}"] = log
            else:   # 10^4..10^5 incompressible doubles: keep the head and a digest of the whole stream
                out[f"draws{i}_head"] = log[:64]
                out[f"draws{i}_count"] = np.array(log.shape[0])
                out[f"draws{i}_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(log).tobytes()).digest(),
                                                        dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, f"sparse_{name}.npz"), **out)
        print("sparse", name, [m.nnz for m in pooled])

    if only:
        return

    # ---------------- row slices of huge graphs (64-bit sort keys; slice-local traces) -----
    # One reference worker (sparse_sampler.py:26-56) run in-process on a contiguous chunk of start nodes
    # of a ring with 2^23 / 2^25 nodes: node ids need 23 / 25 bits, so (node, walk) sort keys exceed 32
    # bits -- the walker's 64-bit-key instantiations -- and the trace covers the slice only.  The ring's
    # Laplacian is translation invariant (two distinct values), so the fixture stores those two values
    # (checked against the reference's get_normalized_laplacian on a 2^12 ring and, for 2^23, on the
    # full graph) and the test rebuilds the CSR procedurally.
    small = lap_sparse(_ring_csr(1 << 12, 1.0, 1.0, adjacency=True)).tocsr()
    diag_v, off_v = float(small[5, 5]), float(small[5, 6])
    assert np.all(small.diagonal() == diag_v) and set(np.unique(small.data)) == {diag_v, off_v}
    for name, log2n, lo_off, n_rows, W, p, L, seed in [
        ("ring2p25_slice_W100", 25, -5, 5, 100, 0.1, 5, 77),     # warp variant, 4 keys per lane, 64-bit keys
        ("ring2p23_slice_W300", 23, -3, 3, 300, 0.1, 4, 78),     # CTA-per-node variant, 64-bit keys
        ("ring2p20_slice_W100", 20, 1000, 6, 100, 0.1, 5, 79),   # 32-bit keys, slice in the middle
    ]:
        n = 1 << log2n
        lap = _ring_csr(n, diag_v, off_v)
        if log2n <= 23:
            ref_lap = lap_sparse(_ring_csr(n, 1.0, 1.0, adjacency=True)).tocsr()
            ref_lap.sort_indices()
            assert np.array_equal(ref_lap.indptr, lap.indptr) and np.array_equal(ref_lap.indices, lap.indices)
            assert np.array_equal(ref_lap.data, lap.data), name
        lo = (n + lo_off) if lo_off < 0 else lo_off
        hi = lo + n_rows
        chunk = list(range(lo, hi))
        ss._init_worker(lap.indptr, lap.indices, lap.data.astype(float, copy=False), n)
        log = []
        real_rng = np.random.default_rng
        np.random.default_rng = lambda sd, _log=log: _RecordingRNG(real_rng(sd), _log)
        try:
            res = ss._worker_walks((chunk, W, p, L, seed, False))
        finally:
            np.random.default_rng = real_rng
        out = {"log2_n": log2n, "lo": lo, "hi": hi, "W": W, "p_halt": p, "L": L, "worker_seed": seed,
               "lap_diag": diag_v, "lap_off": off_v}
        for st in range(L):
            keys = list(res[st].keys())
            m = sp.csr_matrix((np.array([res[st][k] for k in keys], dtype=float),
                               (np.array([k[0] for k in keys], dtype=np.int32),
                                np.array([k[1] for k in keys], dtype=np.int32))), shape=(n, n)) / W   # :125-130
            rows = m[lo:hi].tocsr()
            out[f"step{st}_indptr"] = rows.indptr.astype(np.int64)
            out[f"step{st}_indices"] = rows.indices.astype(np.int64)
            out[f"step{st}_data"] = rows.data.astype(np.float64)
        log = np.array(log, dtype=np.float64).reshape(-1, 2)
        out["draws_head"] = log[:64]
        out["draws_count"] = np.array(log.shape[0])
        out["draws_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(log).tobytes()).digest(), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, f"slice_{name}.npz"), **out)
        print("slice", name, [int(out[f"step{st}_indptr"][-1]) for st in range(L)], "draws", log.shape[0])
        del lap

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

    # ---------------- matvec layer: SparseLinearOperator (M1) and GraphPreprocessor (P1) ----------
    # The real utils_sparse/sparse_lo.py:4-25 and preprocessor/graph_preprocessor.py:85-139 on torch-CPU
    # sparse CSR (the stub LinearOperator base only stores nothing): conversion to float32 / int64, the
    # per-length products M_l @ X and M_l^T @ X (transposed through .t().to_sparse_csr() as the reference
    # does), and the reference's op sequence for one kernel matvec  sum_l f_l M_l (sum_l' f_l' M_l'^T V)
    # (sparse_grf_kernel.py:51-62: ConstantMul + Sum of the SparseLinearOperators) in float32.
    import torch
    from efficient_graph_gp_sparse.preprocessor.graph_preprocessor import GraphPreprocessor as RefPP
    from efficient_graph_gp_sparse.utils_sparse.sparse_lo import SparseLinearOperator as RefLO

    torch.manual_seed(0)
    adj = (_grid(7, 5) + _random_graph(35, 20, 3, weighted=True)).tocsr()
    adj.sum_duplicates()
    adj.sort_indices()
    W, p, L, seed, nproc = 25, 0.1, 4, 9, 3
    pp = RefPP(adj, walks_per_node=W, p_halt=p, max_walk_length=L, random_walk_seed=seed, use_tqdm=False,
               n_processes=nproc)
    ops = pp.preprocess_graph(save_to_disk=False)
    lap = lap_sparse(adj)
    logs, _ = _record_sparse(ss, lap, W, p, L, seed, nproc)
    n = adj.shape[0]
    t = 6
    X = torch.randn(n, t)
    V = torch.randn(n, t)
    f = torch.randn(L)
    mv = {"W": W, "p_halt": p, "L": L, "seed": seed, "n_processes": nproc, "X": X.numpy(), "V": V.numpy(),
          "f": f.numpy()}
    _csr_pack("adj", adj, mv)
    for i, log in enumerate(logs):
        mv[f"draws{i}"] = log
    u = torch.zeros(n, t)
    for st, (m, op) in enumerate(zip(pp.step_matrices_scipy, ops)):
        assert isinstance(op, RefLO) and tuple(op._size()) == (n, n)
        _csr_pack(f"step{st}", m, mv)
        csr = op.sparse_csr_tensor
        mv[f"torch{st}_crow"] = csr.crow_indices().numpy()
        mv[f"torch{st}_col"] = csr.col_indices().numpy()
        mv[f"torch{st}_val"] = csr.values().numpy()                      # float32
        mv[f"matmul{st}"] = op._matmul(X).numpy()
        opt = op._transpose_nonbatch()
        mv[f"tmatmul{st}"] = opt._matmul(X).numpy()
        tc = opt.sparse_csr_tensor
        mv[f"torchT{st}_crow"] = tc.crow_indices().numpy()
        mv[f"torchT{st}_col"] = tc.col_indices().numpy()
        mv[f"torchT{st}_val"] = tc.values().numpy()
        u = u + f[st] * opt._matmul(V)                                   # Phi^T V, length by length
    out = torch.zeros(n, t)
    for st, op in enumerate(ops):
        out = out + f[st] * op._matmul(u)                                # Phi (Phi^T V)
    mv["phiT_V"] = u.numpy()
    mv["K_V"] = out.numpy()
    x1 = torch.tensor([3, 0, 17, 17, 30])
    x2 = torch.tensor([1, 2, 3, 20, 34, 8])
    V2 = torch.randn(x2.numel(), t)
    full = torch.zeros(n, t).index_add_(0, x2, V2)                       # Phi[x2]^T V2 = Phi^T scatter(V2)
    u2 = sum(f[st] * ops[st]._transpose_nonbatch()._matmul(full) for st in range(L))
    mv["x1"], mv["x2"], mv["V2"] = x1.numpy(), x2.numpy(), V2.numpy()
    mv["K_x1x2_V2"] = sum(f[st] * ops[st]._matmul(u2) for st in range(L))[x1].numpy()
    np.savez_compressed(os.path.join(HERE, "matvec_layer.npz"), **mv)
    print("matvec layer written:", [m.nnz for m in pp.step_matrices_scipy])


if __name__ == "__main__":
    main()
