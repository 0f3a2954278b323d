"""Synthetic input graphs of the named benchmark shapes, generated on the device.

Input synthesis only (bench.py, profiles/, tests): nothing here is on the GRF path.  The
reference builds its graphs on the host (networkx / scipy, e.g. scalable_bo/bo_utils/data_utils.py:56-60,
run_scaling_experiment.py:66,160-161); at BASELINE config 4 the host R-MAT generation takes 67 s
and the scipy Laplacian 4 s, against 0.05 s for the Phi build they feed, so the 4 M-node graph is
drawn with torch on the GPU (a second or two) and normalised by grf_laplacian_*.

The draws come from torch's CUDA Philox generator: the same (scale, edges, seed) gives the same
graph on every B200 rank, and ``DeviceGraph.to_scipy()`` hands it to the CPU oracle / reference arm.
"""

from __future__ import annotations

from typing import Tuple

import torch

from .engine import DeviceGraph, _device

RMAT_ABCD = (0.57, 0.19, 0.19, 0.05)      # Graph500 / SURVEY 8d cfg4


def _rmat_pairs(scale: int, m: int, gen: torch.Generator, dev, abcd=RMAT_ABCD) -> torch.Tensor:
    """m R-MAT edge draws as int64 keys min(u, v) * n + max(u, v); self-loops dropped."""
    a, b, c, _ = abcd
    n = 1 << scale
    src = torch.zeros(m, dtype=torch.int64, device=dev)
    dst = torch.zeros(m, dtype=torch.int64, device=dev)
    for bit in range(scale):
        r = torch.rand(m, device=dev, generator=gen)
        src |= (r >= a + b).to(torch.int64) << bit                                   # quadrants c, d
        dst |= (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64) << bit   # quadrants b, d
        del r
    keep = src != dst
    lo = torch.minimum(src, dst)[keep]
    hi = torch.maximum(src, dst)[keep]
    return lo * n + hi


def rmat_adjacency(scale: int, n_edges: int, seed: int = 0, device=None, abcd=RMAT_ABCD,
                   exact: bool = False) -> DeviceGraph:
    """Symmetric unit-weight adjacency (CSR on the device) of an R-MAT graph with 2**scale nodes and at
    least ``n_edges`` undirected edges after symmetrising, de-duplicating and dropping self-loops
    (``exact``: a seeded random subset of exactly ``n_edges`` of them)."""
    dev = _device(device)
    n = 1 << scale
    if n >= 2 ** 31 or 2 * n_edges >= 2 ** 31:
        raise ValueError("graph exceeds int32 index range")
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    want = int(n_edges)
    draw = int(want * 1.06) + 1024
    for _ in range(32):
        keys = torch.unique(torch.cat([keys, _rmat_pairs(scale, draw, gen, dev, abcd)]))
        if keys.numel() >= want:
            break
        draw = int((want - keys.numel()) * 1.25) + 1024
    else:
        raise RuntimeError("R-MAT generator did not reach the requested edge count")
    if exact and keys.numel() > want:
        pick = torch.randperm(keys.numel(), device=dev, generator=gen)[:want]
        keys = keys[pick.sort().values]
    lo, hi = keys // n, keys % n
    del keys
    # symmetrise: entries (lo, hi) and (hi, lo), sorted by (row, col)
    flat = torch.cat([lo * n + hi, hi * n + lo]).sort().values
    del lo, hi
    rows, cols = flat // n, (flat % n).to(torch.int32)
    del flat
    counts = torch.bincount(rows, minlength=n)
    del rows
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    val = torch.ones(cols.numel(), dtype=torch.float64, device=dev)
    return DeviceGraph._from_device(row_ptr.to(torch.int32), cols.contiguous(), val, n)


def rmat_walk_graph(scale: int, n_edges: int, seed: int = 0, device=None) -> Tuple[DeviceGraph, dict]:
    """(normalized Laplacian as the walk graph, statistics) of the R-MAT adjacency above."""
    adj = rmat_adjacency(scale, n_edges, seed, device)
    deg = adj.row_ptr[1:] - adj.row_ptr[:-1]
    stats = {"n_nodes": adj.n_nodes, "undirected_edges": adj.nnz // 2, "max_degree": int(deg.max()),
             "isolated_nodes": int((deg == 0).sum())}
    lap = adj.laplacian()
    stats["nnz_laplacian"] = lap.nnz
    return lap, stats
