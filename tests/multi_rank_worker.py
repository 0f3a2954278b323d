"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.

Every rank builds its row shard of Phi on a power-law graph, multiplies through every exchange the sharded
matvec has -- grf_exchange_sum over peer memory, NCCL all-reduce of all of U, NCCL on the shared columns only
(graph-derived hint and the touched-by-two census) -- and compares its rows with (a) the single-GPU product of
the whole Phi built on the same device and (b) the float64 oracle (rank 0).  Also the row-sharded CG solve.
Prints MULTI_RANK_OK on success; any mismatch raises."""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200"), os.path.join(ROOT, "tests")]

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from grf_b200 import engine, sharding, synth
    from gpu_util import grid_graph
    from oracle import grf_oracle

    # replicated graph, uploaded once per box: every rank one slice + NCCL all-gather == every rank the whole thing
    import scipy.sparse as sp

    adj = sp.random(3000, 3000, 0.01, format="csr", random_state=7)
    adj = ((adj + adj.T) > 0).astype(float).tocsr()
    old_min, engine._UPLOAD_MIN = engine._UPLOAD_MIN, 1 << 10
    try:
        shared = engine.DeviceGraph.laplacian_of(adj, dev, group=True)
    finally:
        engine._UPLOAD_MIN = old_min
    alone = engine.DeviceGraph.laplacian_of(adj, dev)
    for a, b in ((shared.row_ptr, alone.row_ptr), (shared.col_idx, alone.col_idx), (shared.val, alone.val)):
        assert torch.equal(a, b), "slice-upload + all-gather differs from the plain upload"

    t, L, W = 16, 4, 24
    cases = []
    g_pl, _ = synth.rmat_walk_graph(13, 60_000, seed=3, device=dev)                 # power-law: every column shared
    cases.append(("rmat", g_pl, sharding.balanced_bounds(g_pl, world)))
    lap = grf_oracle.normalized_laplacian_sparse(grid_graph(64, 48))                # banded: few shared columns
    g_grid = engine.DeviceGraph.from_scipy(lap, dev)
    cases.append(("grid", g_grid, [int(b) for b in sharding.shard_bounds(g_grid.n_nodes, world)]))
    for name, graph, bounds in cases:
        n = graph.n_nodes
        lo, hi = bounds[rank], bounds[rank + 1]
        cfg = engine.WalkConfig(W, 0.1, L, seed=5)
        torch.manual_seed(0)
        f = torch.randn(L, device=dev)
        v_all = torch.randn(n, t, device=dev)                    # same on every rank (same seed, same device type)
        full = engine.build_phi_blocks(graph, cfg)
        want = full.plan(f, t, merged=False)(v_all)
        part = engine.build_phi_blocks(graph, cfg, lo, hi)
        scale = float(want.abs().max())
        v = v_all[lo:hi].contiguous()
        results = {}
        ex = sharding.make_exchange(n, t, dev, True, mode="peer")
        assert ex.mode == "peer", ex.describe()
        plan = part.plan(f, t, group=True, merged=False, exchange=ex)
        for rep in range(3):                                     # epochs 1..3: the flags are never reset
            results[f"peer{rep}"] = plan(v).clone()
        # all copies of U are bit-identical (rank-ordered sums)
        u0 = ex.u.clone()
        dist.broadcast(u0, src=0)
        assert torch.equal(u0, ex.u), f"{name}: U differs between ranks after grf_exchange_sum"
        results["nccl_all"] = part.plan(f, t, group=True, merged=False,
                                        exchange=sharding.make_exchange(n, t, dev, True, mode="nccl"))(v).clone()
        part.shared_hint = None
        results["nccl_census"] = part.plan(f, t, group=True, merged=False)(v).clone()       # touched-by-two census
        part.shared_hint = graph.shared_columns(bounds, L)
        results["nccl_hint"] = part.plan(f, t, group=True, merged=False)(v).clone()
        if len(part.tblocks) == 1:
            results["nccl_merged"] = part.plan(f, t, group=True, merged=True)(v).clone()
        for key, got in results.items():
            err = float((got - want[lo:hi]).abs().max()) / scale
            assert err <= 2e-5, f"{name}/{key}: rank {rank} differs from the single-GPU product by {err:.2e}"
        if rank == 0:
            mats = [m.astype(np.float32) for m in full.to_scipy_steps()]
            ref = grf_oracle.phi_matvec_f64(mats, f.cpu().numpy(), v_all.cpu().numpy())
            err = np.abs(want.cpu().numpy() - ref).max() / np.abs(ref).max()
            assert err <= 2e-5, f"{name}: single-GPU product vs float64 oracle {err:.2e}"
        # row-sharded CG: (K + 0.5 I) x = b with the sharded matvec == the single-GPU solve
        from grf_b200.cg import linear_cg

        b_all = torch.randn(n, 4, device=dev)
        single = full.plan(f, 4, merged=False)
        # a well-conditioned system (shift ~ the largest row sum of |K 1|): both solvers converge in a few
        # iterations, so the comparison tests the sharded arithmetic and not fp32 CG on a condition number of 10^6
        s2 = float(single(torch.ones(n, 4, device=dev)).abs().max())
        x_want = linear_cg(lambda z: single(z) + s2 * z, b_all, tolerance=1e-6, max_iter=100)
        ex4 = sharding.make_exchange(n, 4, dev, True)
        sharded = part.plan(f, 4, group=True, merged=False, exchange=ex4)
        x_got, info = sharding.sharded_cg(lambda z: sharded(z) + s2 * z, b_all[lo:hi].contiguous(), tolerance=1e-6,
                                          max_iter=100, return_info=True)
        err = float((x_got - x_want[lo:hi]).abs().max()) / float(x_want.abs().max())
        assert err <= 1e-4, f"{name}: CG(sharded) vs single GPU {err:.2e} after {info}"
        from grf_b200.cg import linear_cg_fused

        x_fused, info = linear_cg_fused(sharded, b_all[lo:hi].contiguous(), sigma2=s2, tolerance=1e-6, max_iter=100,
                                        return_info=True)
        err = float((x_fused - x_want[lo:hi]).abs().max()) / float(x_want.abs().max())
        assert err <= 1e-4, f"{name}: CG(fused, sharded) vs single GPU {err:.2e} after {info}"
        dist.barrier()
    if rank == 0:
        print("MULTI_RANK_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
