"""numpy / pure-Python restatement of the reference's GRF hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference file:line it follows; nothing here is imported by the product path.

Reference = MatthewZhang473/Efficient-Gaussian-Process-on-Graphs, paths below
relative to its root.

The walk loop is restated once (:func:`walk_accumulate`) and parameterised by
a *draw source*, because the only thing that differs between "the reference
with its PCG64 stream", "a replayed trace" and "the native counter-based
Philox stream of the CUDA walker" is where the two random numbers of a
continued step come from:

* :class:`GeneratorDraws` -- consumes a ``numpy.random.Generator`` in exactly
  the reference's order (``rng.random()`` then ``rng.integers(deg)``;
  ``sparse_sampler.py:47-51``, ``sampler.py:53-56``) and can record the trace.
* :class:`TraceDraws`     -- replays a recorded ``(trace_u, trace_k)`` pair.
* :class:`PhiloxDraws`    -- Philox4x32-10 keyed by (seed), counter
  (walk id, step // 2); the published algorithm of Salmon et al., SC'11
  (Random123), restated in :func:`philox4x32_10` and pinned against the
  Random123 known-answer vectors in ``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

import math
from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

# load-update semantics of the three walk loops in the reference
LOAD_CUMULATIVE = 0   # load *= deg*w/(1-p)   sparse_sampler.py:54, sampler.py:58
LOAD_LAST_STEP = 1    # load  = deg*w/(1-p)   sampler.py:183 (_sequential_walks)
LOAD_ABLATION = 2     # load  = w             sampler.py:180-181

# how the accumulated sums become matrix entries
SCALE_MUL_RECIP = 0   # csr / W  ==  csr * (1/W)   sparse_sampler.py:130 (scipy _divide)
SCALE_DIV = 1         # value / W                 sampler.py:201


# --------------------------------------------------------------------------
# Laplacians
# --------------------------------------------------------------------------
def normalized_laplacian_sparse(adj) -> sp.csr_matrix:
    """``D^-1/2 (D - A) D^-1/2`` exactly as
    ``efficient_graph_gp_sparse/utils_sparse/graph_utils.py:5-30`` builds it
    (two scipy SpGEMMs, zero-degree rows become empty)."""
    a = adj.tocsr()
    deg = np.array(a.sum(axis=1)).flatten()
    with np.errstate(divide="ignore"):
        dis = 1.0 / np.sqrt(deg)
    dis[np.isinf(dis)] = 0
    d_mat = sp.diags(deg, format="csr")
    dis_mat = sp.diags(dis, format="csr")
    lap = d_mat - a
    return dis_mat @ lap @ dis_mat


def normalized_laplacian_dense(w: np.ndarray) -> np.ndarray:
    """``I - D^-1/2 W D^-1/2`` as the dense branch of
    ``efficient_graph_gp/graph_kernels/utils.py:6-28`` (isolated node keeps a
    diagonal 1, i.e. a self-loop of weight 1)."""
    deg = np.sum(w, axis=1)
    dis = np.zeros_like(w, dtype=float)
    ok = deg > 0
    dis[ok, ok] = 1.0 / np.sqrt(deg[ok])
    return np.eye(w.shape[0]) - dis @ w @ dis


# --------------------------------------------------------------------------
# Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
# as 1, 2, 3", SC'11).  Vectorised over leading axes.
# --------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint64(0x9E3779B9)
_PHILOX_W1 = np.uint64(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)
_SH32 = np.uint64(32)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """``ctr`` (..., 4) uint32, ``key`` (..., 2) uint32 -> (..., 4) uint32."""
    c = np.asarray(ctr, dtype=np.uint64).copy()
    k = np.broadcast_to(np.asarray(key, dtype=np.uint64), c.shape[:-1] + (2,)).copy()
    for _ in range(10):
        p0 = _PHILOX_M0 * c[..., 0]
        p1 = _PHILOX_M1 * c[..., 2]
        hi0, lo0 = p0 >> _SH32, p0 & _MASK32
        hi1, lo1 = p1 >> _SH32, p1 & _MASK32
        n0 = hi1 ^ c[..., 1] ^ k[..., 0]
        n2 = hi0 ^ c[..., 3] ^ k[..., 1]
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k = np.stack([(k[..., 0] + _PHILOX_W0) & _MASK32, (k[..., 1] + _PHILOX_W1) & _MASK32], axis=-1)
    return c.astype(np.uint32)


def halt_threshold(p_halt: float) -> int:
    """Native-mode halting rule: halt iff ``r0 < floor(p_halt * 2**32)``
    (compared as 64-bit integers so that p_halt = 1 halts always)."""
    t = int(math.floor(float(p_halt) * 4294967296.0))
    return max(0, min(t, 1 << 32))


# --------------------------------------------------------------------------
# draw sources
# --------------------------------------------------------------------------
class GeneratorDraws:
    """Sequential numpy Generator, consumed in the reference's order.

    ``record=(trace_u, trace_k)`` stores every draw at ``[walk_id*L + step]``
    so that the CUDA walker can replay it."""

    def __init__(self, rng: np.random.Generator, record=None, dense_choice: bool = False, walk_base: int = 0):
        self.rng = rng
        self.record = record
        self.dense_choice = dense_choice
        self.walk_base = walk_base      # the record holds walks walk_base, walk_base + 1, ... (a row slice)

    def halt(self, walk_id: int, step: int, L: int, p_halt: float) -> bool:
        u = self.rng.random()
        if self.record is not None:
            self.record[0][(walk_id - self.walk_base) * L + step] = u
        return u < p_halt

    def pick(self, walk_id: int, step: int, L: int, deg: int) -> int:
        # rng.choice(neighbors) (sampler.py:56) consumes the stream exactly
        # like rng.integers(deg) (sparse_sampler.py:51); checked in the tests.
        k = int(self.rng.integers(deg))
        if self.record is not None:
            self.record[1][(walk_id - self.walk_base) * L + step] = k
        return k


class TraceDraws:
    def __init__(self, trace_u: np.ndarray, trace_k: np.ndarray, walk_base: int = 0):
        self.u = trace_u
        self.k = trace_k
        self.walk_base = walk_base

    def halt(self, walk_id, step, L, p_halt):
        return self.u[(walk_id - self.walk_base) * L + step] < p_halt

    def pick(self, walk_id, step, L, deg):
        return int(self.k[(walk_id - self.walk_base) * L + step])


class PhiloxDraws:
    """Native stream of the CUDA walker: one Philox4x32-10 block per walk and PAIR of steps:
    counter = (walk_lo, walk_hi, step // 2, 0), key = (seed_lo, seed_hi); words (0, 1) serve
    the even step, words (2, 3) the odd one: the first decides halting
    (``< floor(p_halt * 2**32)``), the second picks the neighbour ``(word * deg) >> 32``."""

    def __init__(self, seed: int):
        self.key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
        self._cache_key = None
        self._cache = None

    def _block(self, walk_id, step):
        ck = (walk_id, step >> 1)
        if self._cache_key != ck:
            ctr = np.array([walk_id & 0xFFFFFFFF, (walk_id >> 32) & 0xFFFFFFFF, step >> 1, 0], dtype=np.uint32)
            self._cache = [int(x) for x in philox4x32_10(ctr, self.key)]
            self._cache_key = ck
        return self._cache

    def halt(self, walk_id, step, L, p_halt):
        return self._block(walk_id, step)[2 * (step & 1)] < halt_threshold(p_halt)

    def pick(self, walk_id, step, L, deg):
        return (self._block(walk_id, step)[2 * (step & 1) + 1] * int(deg)) >> 32


# --------------------------------------------------------------------------
# the walk + accumulate loop
# --------------------------------------------------------------------------
def walk_accumulate(
    indptr: np.ndarray,
    indices: np.ndarray,
    data: np.ndarray,
    starts: Sequence[int],
    num_walks: int,
    p_halt: float,
    max_walk_length: int,
    draws,
    load_mode: int = LOAD_CUMULATIVE,
    accs: Optional[List[Dict[Tuple[int, int], float]]] = None,
) -> List[Dict[Tuple[int, int], float]]:
    """The hot loop of ``sparse_sampler.py:36-54`` (== ``sampler.py:40-59`` on
    a CSR view of the dense matrix; ``sampler.py:163-184`` for the other load
    modes).  Per (step, start, node) the loads are summed in walk order into a
    float64, starting from 0.0, which is what ``defaultdict(float)`` does."""
    L = max_walk_length
    if accs is None:
        accs = [defaultdict(float) for _ in range(L)]
    one_minus_p = 1 - p_halt
    for start in starts:
        start = int(start)
        for w in range(num_walks):
            walk_id = start * num_walks + w
            cur = start
            load = 1.0
            for step in range(L):
                accs[step][(start, cur)] += load
                s = int(indptr[cur])
                deg = int(indptr[cur + 1]) - s
                if deg == 0 or draws.halt(walk_id, step, L, p_halt):
                    break
                k = draws.pick(walk_id, step, L, deg)
                weight = np.float64(data[s + k])
                if load_mode == LOAD_CUMULATIVE:
                    load = load * (np.float64(deg) * weight / one_minus_p)
                elif load_mode == LOAD_LAST_STEP:
                    load = np.float64(deg) * weight / one_minus_p
                else:
                    load = weight
                cur = int(indices[s + k])
    return accs


def accumulators_to_csr(accs, n: int, num_walks: int) -> List[sp.csr_matrix]:
    """``sparse_sampler.py:117-130``: COO -> CSR (sorted indices), then
    ``/ num_walks`` which scipy evaluates as ``* (1 / num_walks)``."""
    mats = []
    for acc in accs:
        if not acc:
            mats.append(sp.csr_matrix((n, n)))
            continue
        keys = list(acc.keys())
        rows = np.array([k[0] for k in keys], dtype=np.int32)
        cols = np.array([k[1] for k in keys], dtype=np.int32)
        vals = np.array([acc[k] for k in keys], dtype=float)
        mats.append(sp.csr_matrix((vals, (rows, cols)), shape=(n, n)) / num_walks)
    return mats


def accumulators_to_dense(accs, n: int, num_walks: int) -> np.ndarray:
    """``sampler.py:188-203``: (N, N, L) zeros, ``value / num_walks``."""
    out = np.zeros((n, n, len(accs)), dtype=float)
    for step, acc in enumerate(accs):
        for (i, j), v in acc.items():
            out[i, j, step] = v / num_walks
    return out


def effective_seed(seed) -> int:
    """``self.seed = seed or 42`` (sparse_sampler.py:65, sampler.py:90)."""
    return seed or 42


def sparse_step_matrices(
    adj,
    num_walks: int,
    p_halt: float,
    max_walk_length: int,
    seed=None,
    n_processes: int = 8,
    record: bool = False,
):
    """``SparseRandomWalk(adj, seed).get_random_walk_matrices(...)`` with the
    reference's start-node sharding (``np.array_split`` over ``n_processes``
    workers, worker i seeded ``seed + i``; sparse_sampler.py:90-107), run
    in-process.  Returns the list of CSR step matrices and, if ``record``,
    the (trace_u, trace_k) pair of the PCG64 draws."""
    a = adj.tocsr()
    n = a.shape[0]
    indptr, indices, data = a.indptr, a.indices, a.data.astype(float, copy=False)
    L = max_walk_length
    trace = None
    if record:
        trace = (np.full(n * num_walks * L, np.nan), np.full(n * num_walks * L, -1, dtype=np.int32))
    accs = [defaultdict(float) for _ in range(L)]
    base = effective_seed(seed)
    for i, chunk in enumerate(np.array_split(np.arange(n), n_processes)):
        draws = GeneratorDraws(np.random.default_rng(base + i), record=trace)
        walk_accumulate(indptr, indices, data, chunk.tolist(), num_walks, p_halt, L, draws, accs=accs)
    mats = accumulators_to_csr(accs, n, num_walks)
    return (mats, trace) if record else mats


def worker_slice_rows(indptr, indices, data, n: int, lo: int, hi: int, num_walks: int, p_halt: float,
                      max_walk_length: int, worker_seed: int, record: bool = False, draws=None):
    """ONE reference worker (``_worker_walks``, sparse_sampler.py:26-56) on the contiguous chunk of start
    nodes ``lo .. hi-1`` with ``default_rng(worker_seed)``, and the rows ``lo .. hi-1`` of the step matrices
    it contributes (``csr_matrix(...) / num_walks``, :125-130) as ``(hi - lo) x n`` CSR.  The trace, if
    recorded, is slice-local: element ``[(walk_id - lo * W) * L + step]``.  Works on graphs far too large
    to walk as a whole (the 2^25-node ring of the 64-bit-key fixtures)."""
    L = max_walk_length
    trace = None
    if record:
        m = (hi - lo) * num_walks * L
        trace = (np.full(m, np.nan), np.full(m, -1, dtype=np.int32))
    if draws is None:
        draws = GeneratorDraws(np.random.default_rng(worker_seed), record=trace, walk_base=lo * num_walks)
    accs = walk_accumulate(indptr, indices, np.asarray(data, dtype=float), range(lo, hi), num_walks, p_halt, L, draws)
    mats = []
    for acc in accs:
        if not acc:
            mats.append(sp.csr_matrix((hi - lo, n)))
            continue
        keys = list(acc.keys())
        rows = np.array([k[0] - lo for k in keys], dtype=np.int32)
        cols = np.array([k[1] for k in keys], dtype=np.int32)
        vals = np.array([acc[k] for k in keys], dtype=float)
        mats.append(sp.csr_matrix((vals, (rows, cols)), shape=(hi - lo, n)) / num_walks)
    return (mats, trace) if record else mats


def ring_laplacian_csr(n: int, diag_v: float, off_v: float) -> sp.csr_matrix:
    """Normalized Laplacian of the unit-weight n-ring built directly (three entries per row, sorted columns);
    the two values come from the reference's get_normalized_laplacian (stored in the slice fixtures)."""
    i = np.arange(n, dtype=np.int64)
    cols = np.stack([(i - 1) % n, i, (i + 1) % n], axis=1)
    vals = np.stack([np.full(n, off_v), np.full(n, diag_v), np.full(n, off_v)], axis=1)
    order = np.argsort(cols, axis=1, kind="stable")
    cols = np.take_along_axis(cols, order, axis=1)
    vals = np.take_along_axis(vals, order, axis=1)
    return sp.csr_matrix((vals.ravel(), cols.ravel().astype(np.int32), (np.arange(n + 1, dtype=np.int64) * 3).astype(np.int32)),
                         shape=(n, n))


def dense_step_tensor(
    adjacency: np.ndarray,
    num_walks: int,
    p_halt: float,
    max_walk_length: int,
    seed=None,
    n_processes: int = 8,
    ablation: bool = False,
    record: bool = False,
):
    """``RandomWalk(Graph(adjacency), seed).get_random_walk_matrices(...)``
    (sampler.py:93-146), including the fall-back to ``_sequential_walks``
    when ``n_processes == 1 or N < 2*n_processes`` (sampler.py:115-116), which
    uses ``default_rng(seed)`` with the seed *as given* (sampler.py:89) and the
    non-cumulative load (sampler.py:183)."""
    n = adjacency.shape[0]
    a = sp.csr_matrix(adjacency)  # row-major nonzeros == np.flatnonzero(row) order
    a.sort_indices()
    indptr, indices, data = a.indptr, a.indices, a.data.astype(float, copy=False)
    L = max_walk_length
    trace = None
    if record:
        trace = (np.full(n * num_walks * L, np.nan), np.full(n * num_walks * L, -1, dtype=np.int32))
    accs = [defaultdict(float) for _ in range(L)]
    if n_processes == 1 or n < n_processes * 2:
        draws = GeneratorDraws(np.random.default_rng(seed), record=trace)
        mode = LOAD_ABLATION if ablation else LOAD_LAST_STEP
        walk_accumulate(indptr, indices, data, range(n), num_walks, p_halt, L, draws, load_mode=mode, accs=accs)
    else:
        base = effective_seed(seed)
        for i, chunk in enumerate(np.array_split(np.arange(n), n_processes)):
            draws = GeneratorDraws(np.random.default_rng(base + i), record=trace)
            walk_accumulate(indptr, indices, data, chunk.tolist(), num_walks, p_halt, L, draws, accs=accs)
    out = accumulators_to_dense(accs, n, num_walks)
    return (out, trace) if record else out


def step_matrices_from_draws(
    adj_csr: sp.csr_matrix,
    num_walks: int,
    p_halt: float,
    max_walk_length: int,
    draws,
    starts=None,
    load_mode: int = LOAD_CUMULATIVE,
) -> List[sp.csr_matrix]:
    """Step matrices for an arbitrary draw source (trace replay / Philox)."""
    a = adj_csr.tocsr()
    n = a.shape[0]
    starts = range(n) if starts is None else starts
    accs = walk_accumulate(
        a.indptr, a.indices, a.data.astype(float, copy=False), starts, num_walks, p_halt, max_walk_length, draws,
        load_mode=load_mode,
    )
    return accumulators_to_csr(accs, n, num_walks)


# --------------------------------------------------------------------------
# kernel assembly
# --------------------------------------------------------------------------
def grf_kernel_sparse(adj, modulator, walks_per_node=50, p_halt=0.1, max_walk_length=10, n_processes=8):
    """``graph_kernels_sparse/fast_grf_kernel_general.py:20-55``."""
    lap = normalized_laplacian_sparse(adj)
    mats = sparse_step_matrices(lap, walks_per_node, p_halt, max_walk_length, seed=None, n_processes=n_processes)
    n = adj.shape[0]
    phi = sp.csr_matrix((n, n))
    for step, f in enumerate(modulator):
        if step < len(mats):
            phi += f * mats[step]
    return phi @ phi.T


def grf_kernel_dense(adj, modulator, walks_per_node=50, p_halt=0.1, max_walk_length=10, n_processes=8):
    """``graph_kernels/fast_grf_kernel_general.py:11-39``."""
    lap = normalized_laplacian_dense(adj)
    feats = dense_step_tensor(lap, walks_per_node, p_halt, max_walk_length, seed=42, n_processes=n_processes)
    phi = feats @ np.asarray(modulator)
    return phi @ phi.T


def exact_series(matrix: np.ndarray, modulator: Sequence[float]) -> np.ndarray:
    """Exact ``Phi = sum_l f_l A^l`` (powers as in
    ``gpflow_kernels/general_kernel_pofm.py:7-42``) and ``K = Phi Phi^T``;
    the target the GRF estimator is unbiased for."""
    n = matrix.shape[0]
    power = np.eye(n)
    phi = np.zeros((n, n))
    for l, f in enumerate(modulator):
        if l > 0:
            power = power @ matrix
        phi += f * power
    return phi @ phi.T


def compute_fro(first, second, relative=True):
    """``utils.py:34-40``."""
    d = np.linalg.norm(first - second)
    return d / np.linalg.norm(first) if relative else d * d


def diffusion_modulator(length: int, beta: float) -> float:
    """``modulation_functions/diffusion_modulator.py:3-6``."""
    return (-beta) ** length / (2 ** length * math.factorial(length))


# --------------------------------------------------------------------------
# matvec / CG layer (parity UNPINNED against the reference -- see __init__)
# --------------------------------------------------------------------------
def phi_matvec_f64(mats: Sequence[sp.csr_matrix], f: Sequence[float], v: np.ndarray, x1=None, x2=None) -> np.ndarray:
    """``K[x1, x2] @ v`` with ``K = Phi Phi^T``, ``Phi = sum_l f_l M_l``
    (``sparse_grf_kernel.py:24-62``) in float64."""
    phi = None
    for fl, m in zip(f, mats):
        term = float(fl) * m.astype(np.float64)
        phi = term if phi is None else phi + term
    phi = phi.tocsr()
    p1 = phi if x1 is None else phi[np.asarray(x1, dtype=np.int64)]
    p2 = phi if x2 is None else phi[np.asarray(x2, dtype=np.int64)]
    return p1 @ (p2.T @ np.asarray(v, dtype=np.float64))


def torch_csr_f32(m: sp.csr_matrix):
    """``GraphPreprocessor.from_scipy_csr`` (graph_preprocessor.py:117-139): int64 indices, float32 values."""
    import torch

    m = m.tocsr()
    return torch.sparse_csr_tensor(torch.from_numpy(m.indptr).long(), torch.from_numpy(m.indices).long(),
                                   torch.from_numpy(m.data).float(), (m.shape[0], m.shape[1]), dtype=torch.float32)


def phi_matvec_reference_torch(mats: Sequence[sp.csr_matrix], f, v, x1=None, x2=None):
    """The reference's op sequence on torch-CPU: fp32 values / int64 indices
    (``graph_preprocessor.py:117-139``), per-length CSR SpMMs
    (``sparse_lo.py:16-18``), transposes via ``.t().to_sparse_csr()``
    (``sparse_lo.py:23-25``), scale + sum (``sparse_grf_kernel.py:59-61``),
    row selection as scatter / gather (upstream InterpolatedLinearOperator)."""
    import torch

    n = mats[0].shape[0]
    ts = []
    for m in mats:
        m = m.tocsr()
        ts.append(
            torch.sparse_csr_tensor(
                torch.from_numpy(m.indptr).long(), torch.from_numpy(m.indices).long(),
                torch.from_numpy(m.data).float(), (n, n), dtype=torch.float32,
            )
        )
    f = torch.as_tensor(f, dtype=torch.float32)
    v = torch.as_tensor(v, dtype=torch.float32)
    if x2 is not None:
        full = torch.zeros((n, v.shape[1]), dtype=torch.float32)
        full.index_add_(0, torch.as_tensor(x2, dtype=torch.long), v)
        v = full
    u = None
    for fl, t in zip(f, ts):
        term = fl * (t.t().to_sparse_csr().matmul(v))
        u = term if u is None else u + term
    out = None
    for fl, t in zip(f, ts):
        term = fl * t.matmul(u)
        out = term if out is None else out + term
    if x1 is not None:
        out = out[torch.as_tensor(x1, dtype=torch.long)]
    return out.numpy()


def phi_fgrad_f64(mats, f, left: np.ndarray, right: np.ndarray, x1=None, x2=None) -> np.ndarray:
    """d/df_l of ``sum(left * (K[x1,x2] @ right))`` -- what upstream
    ``_bilinear_derivative`` yields for the lazy ``sum_l f_l M_l`` operators
    of ``sparse_grf_kernel.py:51-62``."""
    ms = [m.astype(np.float64).tocsr() for m in mats]
    sel1 = (lambda m: m) if x1 is None else (lambda m: m[np.asarray(x1, dtype=np.int64)])
    sel2 = (lambda m: m) if x2 is None else (lambda m: m[np.asarray(x2, dtype=np.int64)])
    a = [sel1(m).T @ left for m in ms]    # M_l[x1]^T left
    b = [sel2(m).T @ right for m in ms]   # M_l[x2]^T right
    fa = sum(float(fl) * x for fl, x in zip(f, a))
    fb = sum(float(fl) * x for fl, x in zip(f, b))
    return np.array([np.sum(a[l] * fb) + np.sum(fa * b[l]) for l in range(len(ms))])


def linear_cg(matmul, rhs: np.ndarray, tolerance: float = 1e-2, max_iter: int = 1000, eps: float = 1e-10):
    """Batched CG with the stopping rule of upstream
    ``linear_operator.utils.linear_cg`` as the reference calls it
    (``models/sparse_grf_model.py:43``): right-hand sides normalised per
    column, zero initial guess, stop when the mean residual norm drops below
    ``tolerance`` after at least ``min(10, max_iter - 1)`` iterations.
    Restated from the published algorithm (the package is not installed)."""
    rhs = np.asarray(rhs, dtype=np.float64)
    norm = np.linalg.norm(rhs, axis=0, keepdims=True)
    norm = np.where(norm < eps, 1.0, norm)
    b = rhs / norm
    x = np.zeros_like(b)
    r = b.copy()
    d = r.copy()
    rs = np.sum(r * r, axis=0, keepdims=True)
    min_iter = min(10, max_iter - 1)
    for k in range(max_iter):
        ad = matmul(d)
        alpha = rs / np.maximum(np.sum(d * ad, axis=0, keepdims=True), eps)
        x = x + alpha * d
        r = r - alpha * ad
        rs_new = np.sum(r * r, axis=0, keepdims=True)
        if k + 1 >= min_iter and np.mean(np.sqrt(rs_new)) < tolerance:
            break
        d = r + (rs_new / np.maximum(rs, eps)) * d
        rs = rs_new
    return x * norm
