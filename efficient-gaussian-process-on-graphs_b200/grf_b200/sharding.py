"""Start-node sharding over the GPUs of one box, and the one collective of the path.

Walks from different start nodes are independent, so GPU g owns a contiguous block of
start nodes = rows of every M_l (the reference shards the same way over processes,
sparse_sampler.py:90; same ``np.array_split`` boundaries).  The CSR walk graph is
replicated; building Phi needs no communication.  The kernel matvec needs exactly one
exchange: the sum over GPUs of the partials ``U_g = Phi_g[x2]^T V_g`` (N x t), done as an
all-reduce over NCCL/NVLink (gloo on CPU in the tests).
"""

from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch


def shard_bounds(n_nodes: int, world_size: int) -> np.ndarray:
    """Row boundaries [b_0 = 0, ..., b_world = n]; shard g owns [b_g, b_g+1).  Same split as
    ``np.array_split(np.arange(n), world)`` (the first n % world shards get one extra row)."""
    base, extra = divmod(int(n_nodes), int(world_size))
    sizes = np.full(world_size, base, dtype=np.int64)
    sizes[:extra] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


def my_rows(n_nodes: int, world_size: int, rank: int) -> Tuple[int, int]:
    b = shard_bounds(n_nodes, world_size)
    return int(b[rank]), int(b[rank + 1])


def owner_of(ids: torch.Tensor, n_nodes: int, world_size: int) -> torch.Tensor:
    """Rank owning each global row id."""
    bounds = torch.as_tensor(shard_bounds(n_nodes, world_size)[1:], device=ids.device)
    return torch.bucketize(ids.to(torch.int64), bounds, right=True)


def sharded_kernel_matvec(apply_t: Callable, apply: Callable, v_local: torch.Tensor, group=None) -> torch.Tensor:
    """out_g = Phi_g (sum_h Phi_h^T V_h): ``apply_t(v_local) -> U_g [N, t]``, all-reduce, ``apply(U) -> out_g``.

    The two callables are this rank's halves of the product (PhiBlocks.apply_t / apply on a
    GPU; anything with the same contract in the CPU tests)."""
    import torch.distributed as dist

    u = apply_t(v_local).contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(u, group=group)
    return apply(u)


def sharded_dot(a: torch.Tensor, b: torch.Tensor, group=None) -> torch.Tensor:
    """Column-wise dot products of row-sharded [n_g, t] blocks (the t-float all-reduce of a
    row-sharded CG iteration)."""
    import torch.distributed as dist

    d = (a * b).sum(dim=0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(d, group=group)
    return d


def shared_columns(touched: torch.Tensor, group=None, max_fraction: float = 0.5) -> Optional[torch.Tensor]:
    """Columns of Phi touched by at least two row shards (same result on every rank), or None when
    they are more than ``max_fraction`` of all columns (then a plain all-reduce of U is cheaper).

    ``touched``: bool / int [N], this shard's non-empty columns.  A column touched by ONE shard needs
    no communication at all: its sum is that shard's partial, and only that shard reads it in the
    second half of the product.  For a banded Phi the shared columns are just the bands around the
    shard boundaries, so the exchange shrinks from N x t to O(bandwidth x t)."""
    import torch.distributed as dist

    count = touched.to(torch.int32).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(count, group=group)
    else:
        return torch.zeros(0, dtype=torch.int64, device=touched.device)
    shared = torch.nonzero(count >= 2).flatten()
    if shared.numel() > max_fraction * touched.numel():
        return None
    return shared


def reduce_shared(u: torch.Tensor, shared: Optional[torch.Tensor], buf: Optional[torch.Tensor] = None, group=None):
    """Sum the per-shard partials ``u`` [N, ld] over the ranks, exchanging only the shared columns."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return u
    if shared is None:
        dist.all_reduce(u, group=group)
        return u
    if shared.numel() == 0:
        return u
    if buf is None:
        buf = torch.empty((shared.numel(), u.shape[1]), dtype=u.dtype, device=u.device)
    torch.index_select(u, 0, shared, out=buf)
    dist.all_reduce(buf, group=group)
    u.index_copy_(0, shared, buf)
    return u
