"""CPU baseline of the GRF hot path for bench.py (TEST / MEASUREMENT INFRASTRUCTURE ONLY).

The reference is pure Python, so its faithful CPU port is pure Python too:
``sampler_pool`` restates ``SparseRandomWalk.get_random_walk_matrices``
(sparse_sampler.py:72-132) -- fork pool over ``np.array_split`` chunks of start
nodes, one PCG64 ``default_rng(seed + i)`` per worker, ``defaultdict(float)``
accumulators keyed by ``(start, node)``, dict merge in the parent, COO -> CSR
and ``/ num_walks`` -- with the same per-visit work (two numpy scalar RNG
calls, numpy scalar indexing, tuple hashing), so its speed is the reference's
speed.  ``tests/test_oracle_golden.py::test_cpu_baseline_port_matches_reference``
pins its output bit-for-bit to the reference-generated fixtures.

``matvec_torch_cpu`` restates the reference's per-matvec op sequence
(sparse_lo.py:16-25, sparse_grf_kernel.py:59-61) on torch-CPU CSR tensors.
"""

from __future__ import annotations

import multiprocessing as mp
import os
import time
from collections import defaultdict
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import scipy.sparse as sp

_CSR = {}


def _bind(indptr, indices, data):
    _CSR["indptr"], _CSR["indices"], _CSR["data"] = indptr, indices, data


def _chunk_walks(job):
    """One worker: all walks of a chunk of start nodes (sparse_sampler.py:26-56)."""
    starts, num_walks, p_halt, length, seed = job
    indptr, indices, data = _CSR["indptr"], _CSR["indices"], _CSR["data"]
    rng = np.random.default_rng(seed)
    acc = [defaultdict(float) for _ in range(length)]
    keep = 1 - p_halt
    visits = 0
    for start in starts:
        for _ in range(num_walks):
            node, load = start, 1.0
            for step in range(length):
                acc[step][(start, node)] += load
                visits += 1
                lo = indptr[node]
                deg = indptr[node + 1] - lo
                if deg == 0 or rng.random() < p_halt:
                    break
                pick = rng.integers(deg)
                load *= deg * data[lo + pick] / keep
                node = indices[lo + pick]
    return acc, visits


def sampler_pool(adj_csr, num_walks, p_halt, length, seed=None, n_processes=None, starts=None,
                 return_visits=False):
    """Step matrices of the start nodes ``starts`` (default: all), reference algorithm."""
    a = adj_csr.tocsr()
    n = a.shape[0]
    n_processes = n_processes or os.cpu_count()
    base = seed or 42
    nodes = np.arange(n) if starts is None else np.asarray(starts)
    chunks = np.array_split(nodes, n_processes)
    jobs = [(c.tolist(), num_walks, p_halt, length, base + i) for i, c in enumerate(chunks)]
    merged = [defaultdict(float) for _ in range(length)]
    data = a.data.astype(float, copy=False)
    if n_processes == 1:
        _bind(a.indptr, a.indices, data)
        results = [_chunk_walks(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=n_processes, mp_context=mp.get_context("fork"), initializer=_bind,
                                 initargs=(a.indptr, a.indices, data)) as pool:
            results = list(pool.map(_chunk_walks, jobs))
    visits = sum(r[1] for r in results)
    for res, _ in results:
        for step in range(length):
            for key, value in res[step].items():
                merged[step][key] += value
    mats = []
    for step in range(length):
        d = merged[step]
        if not d:
            mats.append(sp.csr_matrix((n, n)))
            continue
        keys = list(d.keys())
        rows = np.fromiter((k[0] for k in keys), dtype=np.int32, count=len(keys))
        cols = np.fromiter((k[1] for k in keys), dtype=np.int32, count=len(keys))
        vals = np.fromiter((d[k] for k in keys), dtype=float, count=len(keys))
        mats.append(sp.csr_matrix((vals, (rows, cols)), shape=(n, n)) / num_walks)
    return (mats, visits) if return_visits else mats


def time_sampler(adj_csr, num_walks, p_halt, length, starts, n_processes=None):
    """(seconds, walk_steps) for one pass of the reference algorithm over ``starts``."""
    t0 = time.perf_counter()
    _, visits = sampler_pool(adj_csr, num_walks, p_halt, length, seed=42, n_processes=n_processes, starts=starts,
                             return_visits=True)
    return time.perf_counter() - t0, visits


def matvec_torch_cpu(mats, f, v, threads=None):
    """Builds the reference's operators once; returns a closure running one Phi(Phi^T V)."""
    import torch

    if threads:
        torch.set_num_threads(threads)
    n = mats[0].shape[0]
    ops, ops_t = [], []
    for m in mats:
        m = m.tocsr()
        t = torch.sparse_csr_tensor(torch.from_numpy(m.indptr).long(), torch.from_numpy(m.indices).long(),
                                    torch.from_numpy(m.data).float(), (n, n), dtype=torch.float32)
        ops.append(t)
    f = torch.as_tensor(f, dtype=torch.float32)
    v = torch.as_tensor(v, dtype=torch.float32)

    def run():
        u = None
        for fl, t in zip(f, ops):
            term = fl * t.t().to_sparse_csr().matmul(v)       # sparse_lo.py:25 re-sorts on every call
            u = term if u is None else u + term
        out = None
        for fl, t in zip(f, ops):
            term = fl * t.matmul(u)
            out = term if out is None else out + term
        return out

    return run
