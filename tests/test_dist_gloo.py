"""The N > 1 path on CPU: world_size-2 gloo, host logic only (shard bounds, the U all-reduce,
row-sharded CG dot products).  The per-rank halves of the product are played by the float64
oracle here; on GPUs they are PhiBlocks.apply_t / apply (tests/test_gpu_matvec.py checks that
those two decompose the same way on one device)."""

import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys

    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from grf_b200 import sharding
    from oracle import c_oracle, grf_oracle as orc

    # the same graph / Phi on every rank (CSR graph replicated), rows sharded
    rng = np.random.default_rng(0)
    n = 101
    i = np.arange(n)
    a = sp.csr_matrix((np.ones(2 * n), (np.r_[i, i], np.r_[(i + 1) % n, (i - 1) % n])), shape=(n, n))   # ring: banded Phi
    lap = orc.normalized_laplacian_sparse(a)
    lo, hi = sharding.my_rows(n, world, rank)
    mats = c_oracle.step_matrices(lap, 12, 0.1, 3, seed=5, start_lo=lo, start_hi=hi)   # this rank's rows only
    f = np.array([1.0, -0.5, 0.25])
    phi_g = sum(fl * m for fl, m in zip(f, mats)).tocsr()                               # (n_g, n)
    v_full = rng.standard_normal((n, 4))
    v_g = torch.tensor(v_full[lo:hi])

    out_g = sharding.sharded_kernel_matvec(
        lambda v: torch.tensor(phi_g.T @ v.numpy()), lambda u: torch.tensor(phi_g @ u.numpy()), v_g)
    dots = sharding.sharded_dot(out_g, v_g)
    # shared-column exchange: only columns touched by both shards travel; every shard ends up with the
    # full sum on the columns IT touches (the only ones its second half reads)
    touched = torch.tensor(np.asarray((phi_g != 0).sum(axis=0)).ravel() > 0)
    shared = sharding.shared_columns(touched, max_fraction=1.0)
    u_g = torch.tensor(phi_g.T @ v_g.numpy())
    u_full = u_g.clone()
    dist.all_reduce(u_full)
    u_sh = sharding.reduce_shared(u_g.clone(), shared)
    assert shared is not None and 0 < shared.numel() < n
    assert torch.allclose(u_sh[touched], u_full[touched], rtol=1e-13, atol=1e-13)
    assert sharding.shared_columns(touched, max_fraction=0.0) is None          # falls back to the full all-reduce
    assert torch.allclose(sharding.reduce_shared(u_g.clone(), None), u_full)
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out_g.numpy())
    np.save(os.path.join(out_dir, f"dots{rank}.npy"), dots.numpy())
    # the Exchange object of the sharded plan in its all-reduce mode (what a box without peer memory runs)
    ex = sharding.make_exchange(n, 4, "cpu", True, mode="auto")
    assert ex.mode == "nccl" and tuple(ex.u.shape) == (n, 4) and "all-reduce" in ex.describe()
    ex.u.copy_(u_g.float())
    assert torch.allclose(ex.reduce().double(), u_full, rtol=1e-5, atol=1e-5)
    # row-sharded CG: (Phi Phi^T + 0.3 I) x = b with all-reduced dot products == the dense solve
    b_full = torch.tensor(rng.standard_normal((n, 3)), dtype=torch.float32)

    def product(z):      # this rank's rows of K z, float32 vectors as on the GPU
        u = torch.tensor(phi_g.T @ z.double().numpy())
        dist.all_reduce(u)
        return torch.tensor(phi_g @ u.numpy(), dtype=torch.float32) + 0.3 * z

    x_g, info = sharding.sharded_cg(product, b_full[lo:hi].clone(), tolerance=1e-7, max_iter=300, return_info=True)
    np.save(os.path.join(out_dir, f"cg{rank}.npy"), x_g.numpy())
    np.save(os.path.join(out_dir, "cg_b.npy"), b_full.numpy())
    dist.destroy_process_group()


def test_shard_bounds_match_array_split():
    from grf_b200 import sharding

    for n in (0, 1, 7, 100, 99856, 4_000_003):
        for w in (1, 2, 3, 8):
            want = [len(c) for c in np.array_split(np.arange(min(n, 50_000)), w)] if n <= 50_000 else None
            b = sharding.shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0) and np.diff(b).max() - np.diff(b).min() <= 1
            if want is not None:
                assert list(np.diff(b)) == want
    ids = torch.tensor([0, 49, 50, 99, 100])
    assert sharding.owner_of(ids, 101, 2).tolist() == [0, 0, 0, 1, 1]


def test_two_rank_sharded_matvec_equals_unsharded(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from grf_b200 import sharding
    from oracle import c_oracle, grf_oracle as orc

    rng = np.random.default_rng(0)
    n = 101
    i = np.arange(n)
    a = sp.csr_matrix((np.ones(2 * n), (np.r_[i, i], np.r_[(i + 1) % n, (i - 1) % n])), shape=(n, n))
    lap = orc.normalized_laplacian_sparse(a)
    mats = c_oracle.step_matrices(lap, 12, 0.1, 3, seed=5)
    v = rng.standard_normal((n, 4))
    want = orc.phi_matvec_f64(mats, [1.0, -0.5, 0.25], v)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)])
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
    dots = [np.load(tmp_path / f"dots{r}.npy") for r in range(world)]
    assert np.allclose(dots[0], dots[1]) and np.allclose(dots[0], (want * v).sum(0))
    phi = sum(fl * m for fl, m in zip([1.0, -0.5, 0.25], mats)).toarray()
    b = np.load(tmp_path / "cg_b.npy").astype(np.float64)
    x_want = np.linalg.solve(phi @ phi.T + 0.3 * np.eye(n), b)
    x_got = np.concatenate([np.load(tmp_path / f"cg{r}.npy") for r in range(world)])
    assert np.abs(x_got - x_want).max() <= 1e-4 * np.abs(x_want).max()


def test_balanced_bounds_equalise_the_work_estimate():
    """Strong-scaling shards: contiguous, monotone, equal sums of the per-node work estimate (isolated start nodes
    count 0.15 of a walking one) -- on a graph whose isolated nodes are all at the high ids."""
    from types import SimpleNamespace

    from grf_b200 import sharding

    deg = torch.cat([torch.full((1000,), 5), torch.zeros(3000, dtype=torch.int64)])
    graph = SimpleNamespace(n_nodes=4000, row_ptr=torch.cat([torch.zeros(1, dtype=torch.int64), deg.cumsum(0)]))
    b = sharding.balanced_bounds(graph, 4)
    assert b[0] == 0 and b[-1] == 4000 and all(x <= y for x, y in zip(b, b[1:]))
    w = torch.where(deg > 0, 1.0, 0.15)
    loads = [float(w[b[i]:b[i + 1]].sum()) for i in range(4)]
    assert max(loads) - min(loads) <= 1.0 + 1e-9
    assert b[1] < 1000 and b[3] > 1000          # not equal counts: the walking nodes are split over several shards
    assert sharding.balanced_bounds(graph, 1) == [0, 4000]
