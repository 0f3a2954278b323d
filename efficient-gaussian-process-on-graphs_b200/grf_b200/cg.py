"""Batched conjugate gradients on the device.

Restates the semantics of upstream ``linear_operator.utils.linear_cg`` as the
reference calls it (``models/sparse_grf_model.py:43``: ``linear_cg(A._matmul,
b.T, tolerance=cg_tolerance)``): right-hand sides normalised per column, zero
initial guess, no preconditioner, stop when the mean residual norm falls below
``tolerance`` after at least ``min(10, max_iter - 1)`` iterations, at most
``max_iter`` (upstream default 1000).  The package itself is not installed
here, so parity is against a float64 direct solve (tests), not upstream.

Every iteration is one kernel matvec (two grf_b200 SpMM launches) plus O(n t)
vector updates; the convergence check reads one scalar back every
``check_every`` iterations instead of every iteration.
"""

from __future__ import annotations

from typing import Callable

import torch


def linear_cg(matmul_closure: Callable[[torch.Tensor], torch.Tensor], rhs: torch.Tensor, tolerance: float = 1e-2,
              max_iter: int = 1000, eps: float = 1e-10, check_every: int = 4, return_info: bool = False):
    squeeze = rhs.dim() == 1
    b = rhs[:, None] if squeeze else rhs
    b = b.to(torch.float32)
    norm = b.norm(dim=0, keepdim=True)
    norm = torch.where(norm < eps, torch.ones_like(norm), norm)
    b = b / norm
    x = torch.zeros_like(b)
    r = b.clone()
    d = r.clone()
    rs = (r * r).sum(dim=0, keepdim=True)
    min_iter = min(10, max_iter - 1)
    iters = 0
    for k in range(max_iter):
        ad = matmul_closure(d)
        alpha = rs / (d * ad).sum(dim=0, keepdim=True).clamp_min(eps)
        x = x + alpha * d
        r = r - alpha * ad
        rs_new = (r * r).sum(dim=0, keepdim=True)
        iters = k + 1
        if iters >= min_iter and (iters % check_every == 0 or iters == max_iter):
            if float(rs_new.sqrt().mean()) < tolerance:
                break
        d = r + (rs_new / rs.clamp_min(eps)) * d
        rs = rs_new
    out = x * norm
    out = out[:, 0] if squeeze else out
    return (out, {"iterations": iters}) if return_info else out


def linear_cg_fused(plan, rhs: torch.Tensor, sigma2: float = 0.0, tolerance: float = 1e-2, max_iter: int = 1000,
                    eps: float = 1e-10, check_every: int = 4, return_info: bool = False, use_graph: bool = True):
    """Solve ``(K + sigma2 I) X = rhs`` with ``K`` = a square :class:`~grf_b200.engine.MatvecPlan`.

    Same iteration and stopping rule as :func:`linear_cg`; every iteration is the two spmm
    launches of the plan plus three fused CUDA launches (``csrc/grf_cg.cu``) instead of a dozen
    elementwise / reduction launches, and the dot products are reduced in a fixed order."""
    import ctypes

    from . import _lib
    from ._lib import check

    if plan.n1 != plan.n2:
        raise ValueError("linear_cg_fused needs a square operator (x1 and x2 of equal length)")
    # a row-sharded plan: the vectors hold this rank's rows; every dot product is one t-float all-reduce
    # of the per-rank partial sums (two per iteration), the rest is the same three fused launches
    group = None
    if plan.group is not None:
        import torch.distributed as dist

        group = (None if plan.group is True else plan.group, dist)
        use_graph = False

    def allsum(partials):
        """[n_part, t] per-rank partial sums -> the same buffer holding the global sum in row 0."""
        if group is None:
            return
        total = partials.sum(dim=0)
        group[1].all_reduce(total, group=group[0])
        partials.zero_()
        partials[0] = total

    L = _lib.lib()
    dev = plan._dev
    squeeze = rhs.dim() == 1
    b = (rhs[:, None] if squeeze else rhs).to(device=dev, dtype=torch.float32)
    n, t = b.shape
    if t != plan.t or n != plan.n2:
        raise ValueError("rhs shape does not match the plan")
    sq = (b * b).sum(dim=0, keepdim=True)
    if group is not None:
        group[1].all_reduce(sq, group=group[0])
    norm = sq.sqrt()
    norm = torch.where(norm < eps, torch.ones_like(norm), norm)
    b = (b / norm).contiguous()
    x = torch.zeros_like(b)
    r = b.clone()
    d = b.clone()
    kd = torch.empty_like(b)
    rs = (r * r).sum(dim=0).contiguous()
    if group is not None:
        group[1].all_reduce(rs, group=group[0])
    rs_next = torch.empty_like(rs)
    n_part = L.grf_cg_num_partials(n, t)
    pa = torch.empty((n_part, t), dtype=torch.float32, device=dev)
    pb = torch.empty((n_part, t), dtype=torch.float32, device=dev)
    def P(tensor):
        return ctypes.c_void_p(tensor.data_ptr())

    min_iter = min(10, max_iter - 1)
    state = {"rs": rs, "rs_next": rs_next}

    def iterate(count):
        for _ in range(count):
            plan(d, kd)
            check(L.grf_cg_dot(P(kd), t, P(d), t, float(sigma2), n, t, P(pa), stream_of()))
            allsum(pa)
            check(L.grf_cg_update(P(x), t, P(r), t, P(d), t, P(kd), t, P(state["rs"]), P(pa), n, t, float(eps),
                                  P(pb), stream_of()))
            allsum(pb)
            check(L.grf_cg_direction(P(d), t, P(r), t, P(state["rs"]), P(pb), n, t, float(eps),
                                     P(state["rs_next"]), stream_of()))
            state["rs"], state["rs_next"] = state["rs_next"], state["rs"]

    def stream_of():
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def converged():
        return float(state["rs"].sqrt().mean()) < tolerance

    iters = 0
    block = max(2, check_every + (check_every & 1))  # even: the rs ping-pong returns to its start
    # capturing costs a few ms: worth it only where launches dominate (small systems, many iterations)
    if use_graph and max_iter >= 4 * block and n * t <= 262144:
        # a CG iteration here is ~5 short launches; replaying a captured block of them removes the
        # Python / launch overhead that otherwise dominates small systems (a BO training set)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            iterate(block)                      # real iterations 1..block (also the warm-up)
        torch.cuda.current_stream(dev).wait_stream(side)
        iters = block
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            iterate(block)                      # captured, not executed
        while iters < max_iter:
            if iters >= min_iter and converged():
                break
            graph.replay()
            iters += block
    else:
        while iters < max_iter:
            step = min(block, max_iter - iters)
            iterate(step)
            iters += step
            if iters >= min_iter and converged():
                break
    out = x * norm
    out = out[:, 0] if squeeze else out
    return (out, {"iterations": iters}) if return_info else out
