"""-m gpu parity tests of the Phi blocks, the transpose and the fused Phi(Phi^T V) matvec.

Tolerance: the matvec computes in fp32 (the reference's dtype on device,
graph_preprocessor.py:131-139); results are compared with a float64 scipy
evaluation of the same fp32-rounded Phi to fp32 round-off:
|got - want| <= 2e-5 * max|want| (stated per test)."""

import numpy as np
import pytest
import scipy.sparse as sp

from gpu_util import grid_graph, powerlaw_graph, random_graph

pytestmark = pytest.mark.gpu

RTOL = 2e-5


@pytest.fixture(scope="module")
def env():
    import torch

    assert torch.cuda.is_available(), "these tests need a B200"
    from grf_b200 import _lib, engine
    from oracle import c_oracle, grf_oracle

    return dict(torch=torch, lib=_lib, eng=engine, c=c_oracle, o=grf_oracle)


@pytest.fixture(scope="module")
def case(env):
    """A Phi built natively on the GPU + the same step matrices on the host."""
    eng = env["eng"]
    lap = env["o"].normalized_laplacian_sparse(random_graph(700, 2400, 5, weighted=True))
    g = eng.DeviceGraph.from_scipy(lap)
    cfg = eng.WalkConfig(40, 0.1, 4, seed=11)
    steps = eng.build_step_matrices(g, cfg)
    phi = eng.build_phi_blocks(g, cfg)
    return dict(lap=lap, steps=steps, mats=steps.to_scipy(), phi=phi, cfg=cfg, g=g)


def _close(got, want, rtol=RTOL):
    scale = max(1e-30, float(np.max(np.abs(want))))
    return float(np.max(np.abs(got - want))) <= rtol * scale


def test_blocks_hold_float32_rounding_of_the_step_matrices(env, case):
    """entries == torch .float() of the float64 step matrices (graph_preprocessor.py:131-139), bit for bit,
    whether built straight from staging or from the reference layout."""
    eng = env["eng"]
    via_steps = eng.PhiBlocks.from_step_matrices(case["steps"])
    for phi in (case["phi"], via_steps):
        back = phi.to_scipy_steps()
        for s, m in enumerate(case["mats"]):
            assert np.array_equal(back[s].indptr, m.indptr) and np.array_equal(back[s].indices, m.indices)
            assert np.array_equal(back[s].data.view(np.int32), m.data.astype(np.float32).view(np.int32))
    assert int(case["phi"].visits) == case["steps"].visits


def test_blocks_from_scipy_and_torch_csr(env, case):
    eng, torch = env["eng"], env["torch"]
    a = eng.phi_blocks_from_scipy(case["mats"])
    ts = []
    for m in case["mats"]:
        ts.append(torch.sparse_csr_tensor(torch.from_numpy(m.indptr).long(), torch.from_numpy(m.indices).long(),
                                          torch.from_numpy(m.data).float(), m.shape).cuda())
    b = eng.phi_blocks_from_torch_csr(ts)
    for phi in (a, b):
        assert torch.equal(phi.blk_ptr, case["phi"].blk_ptr)
        assert torch.equal(phi.entries, case["phi"].entries)


def test_transpose_is_exact_and_row_sorted(env, case):
    phi = case["phi"]
    L, n = phi.n_steps, phi.n_cols
    tptr = phi.tblk_ptr.cpu().numpy().astype(np.int64)
    tent = phi.tentries.cpu().numpy()
    rows, vals = tent[:, 0] & ((1 << 27) - 1), tent[:, 1].copy().view(np.float32)
    steps_of = tent[:, 0] >> 27
    for s, m in enumerate(case["mats"]):
        mt = m.astype(np.float32).T.tocsr()
        mt.sort_indices()
        for c in range(n):
            b, e = tptr[c * L + s], tptr[c * L + s + 1]
            assert np.array_equal(rows[b:e], mt.indices[mt.indptr[c]:mt.indptr[c + 1]])
            assert np.all(steps_of[b:e] == s)
            assert np.array_equal(vals[b:e], mt.data[mt.indptr[c]:mt.indptr[c + 1]])


def test_transpose_with_hub_columns_is_row_sorted(env):
    eng = env["eng"]
    lap = env["o"].normalized_laplacian_sparse(powerlaw_graph(4000, 30000, 2))
    g = eng.DeviceGraph.from_scipy(lap)
    phi = eng.build_phi_blocks(g, eng.WalkConfig(30, 0.1, 3, seed=3))
    tptr = phi.tblk_ptr.cpu().numpy().astype(np.int64)
    seg = np.diff(tptr)
    assert seg.max() > 256 and ((seg > 32) & (seg <= 256)).any(), "test graph should have hub, mid and short segments"
    rows = phi.tentries.cpu().numpy()[:, 0] & ((1 << 27) - 1)
    for g0 in np.flatnonzero(seg > 1):
        r = rows[tptr[g0]:tptr[g0 + 1]]
        assert np.all(r[1:] > r[:-1])
    # and it is the same multiset as Phi
    back = phi.to_scipy_steps()
    assert sum(m.nnz for m in back) == len(rows)


def test_transpose_sorts_segments_longer_than_the_shared_memory_stage(env):
    """A star: every row reaches the hub, so the hub's (column, length) segments hold ~n entries --
    more than the 16384 the long-segment sort stages in shared memory (tiled path), and sizes
    (1024, 4096] / (4096, 16384] appear on the way there."""
    eng = env["eng"]
    for n in (3000, 12000, 40000):
        leaves = np.arange(1, n)
        adj = sp.csr_matrix((np.ones(2 * (n - 1)), (np.r_[np.zeros(n - 1, dtype=np.int64), leaves],
                                                    np.r_[leaves, np.zeros(n - 1, dtype=np.int64)])), shape=(n, n))
        lap = env["o"].normalized_laplacian_sparse(adj)
        g = eng.DeviceGraph.from_scipy(lap)
        phi = eng.build_phi_blocks(g, eng.WalkConfig(8, 0.1, 3, seed=5))
        tptr = phi.tblk_ptr.cpu().numpy().astype(np.int64)
        seg = np.diff(tptr)
        assert seg.max() > 0.5 * n
        rows = phi.tentries.cpu().numpy()[:, 0] & ((1 << 27) - 1)
        for g0 in np.flatnonzero(seg > 1):
            r = rows[tptr[g0]:tptr[g0 + 1]]
            assert np.all(r[1:] > r[:-1]), (n, g0, seg[g0])
        want = sp.vstack([m for m in phi.to_scipy_steps()]).tocsc()      # same multiset, column by column
        got_cols = np.repeat(np.arange(n), np.add.reduceat(seg, np.arange(0, seg.size, 3)))
        assert np.array_equal(np.bincount(got_cols, minlength=n), np.diff(want.indptr))


@pytest.mark.parametrize("t", [1, 2, 3, 4, 8, 12, 16, 17, 32, 64, 65, 130])
def test_matvec_matches_float64(env, case, t):
    torch, o = env["torch"], env["o"]
    phi = case["phi"]
    rng = np.random.default_rng(t)
    f = rng.standard_normal(phi.n_steps)
    v = rng.standard_normal((phi.n_rows, t)).astype(np.float32)
    got = phi.matvec(torch.tensor(f), torch.tensor(v).cuda()).cpu().numpy()
    mats32 = [m.astype(np.float32) for m in case["mats"]]
    want = o.phi_matvec_f64(mats32, f.astype(np.float32), v)
    assert got.shape == (phi.n_rows, t)
    assert _close(got, want)


def test_matvec_vector_rhs_and_reference_op_sequence(env, case):
    """1-D rhs; and agreement with the reference's own op sequence (torch CPU CSR, fp32)."""
    torch, o = env["torch"], env["o"]
    phi = case["phi"]
    rng = np.random.default_rng(0)
    f = rng.standard_normal(phi.n_steps).astype(np.float32)
    v = rng.standard_normal((phi.n_rows, 16)).astype(np.float32)
    got = phi.matvec(torch.tensor(f), torch.tensor(v).cuda()).cpu().numpy()
    ref = o.phi_matvec_reference_torch(case["mats"], f, v)
    assert _close(got, ref, rtol=1e-4)
    g1 = phi.matvec(torch.tensor(f), torch.tensor(v[:, 0]).cuda()).cpu().numpy()
    assert g1.shape == (phi.n_rows,) and _close(g1, got[:, 0], rtol=1e-6)


@pytest.mark.parametrize("t", [1, 16, 17])
def test_matvec_row_subsets(env, case, t):
    """K[x1, x2] @ v with arbitrary (unsorted, repeated) index sets -- sparse_grf_kernel.py:32-41."""
    torch, o = env["torch"], env["o"]
    phi = case["phi"]
    rng = np.random.default_rng(100 + t)
    x1 = rng.integers(0, phi.n_rows, size=211)
    x2 = np.r_[rng.permutation(phi.n_rows)[:300], [5, 5, 17]]       # with repeats
    f = rng.standard_normal(phi.n_steps).astype(np.float32)
    v = rng.standard_normal((x2.size, t)).astype(np.float32)
    mats32 = [m.astype(np.float32) for m in case["mats"]]
    want = o.phi_matvec_f64(mats32, f, v, x1=x1, x2=x2)
    got = phi.matvec(torch.tensor(f), torch.tensor(v).cuda(), x1=torch.tensor(x1).cuda(),
                     x2=torch.tensor(x2).cuda()).cpu().numpy()
    assert got.shape == (211, t) and _close(got, want)
    # x as float column tensors, the way the reference's models pass them (x.long().flatten())
    got2 = phi.matvec(torch.tensor(f), torch.tensor(v).cuda(), x1=torch.tensor(x1, dtype=torch.float32)[:, None].cuda(),
                      x2=torch.tensor(x2, dtype=torch.float32)[:, None].cuda()).cpu().numpy()
    assert _close(got2, got, rtol=1e-6)      # repeated ids are scatter-added with atomics: order may differ


@pytest.mark.parametrize("kind", ["grid", "ring", "periodic_grid"])
@pytest.mark.parametrize("t", [4, 16, 32, 12])
def test_tiled_matvec_matches_gather_matvec(env, kind, t):
    """Banded Phi takes the shared-memory-tiled kernel; it must agree with the global-gather
    kernel and with float64.  The periodic grid has wrap-around edges, so some chunks have a
    window that does not fit and fall back inside the same launch."""
    eng, torch, o = env["eng"], env["torch"], env["o"]
    from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian
    from gpu_util import ring_graph

    if kind == "grid":
        adj = grid_graph(150, 120)
    elif kind == "ring":
        adj = ring_graph(30000)
    else:
        nx, ny = 400, 60          # rows of 400 nodes, periodic in y: node i <-> i + (ny-1)*nx
        adj = grid_graph(nx, ny).tolil()
        for x in range(nx):
            adj[x, (ny - 1) * nx + x] = 1.0
            adj[(ny - 1) * nx + x, x] = 1.0
        adj = adj.tocsr()
    lap = get_normalized_laplacian(adj)
    g = eng.DeviceGraph.from_scipy(lap)
    phi = eng.build_phi_blocks(g, eng.WalkConfig(20, 0.1, 4, seed=8))
    phi.use_tiles = True
    phi.build_windows()
    assert phi.win is not None and phi.win_max_width > 0
    gen = torch.Generator(device="cuda").manual_seed(t)
    f = torch.randn(4, device="cuda", generator=gen)
    v = torch.randn(phi.n_rows, t, device="cuda", generator=gen)
    tiled = phi.matvec(f, v).clone()
    phi.use_tiles = False
    plain = phi.matvec(f, v).clone()
    phi.use_tiles = True
    scale = float(plain.abs().max())
    assert float((tiled - plain).abs().max()) <= 1e-5 * scale
    mats32 = phi.to_scipy_steps()
    want = o.phi_matvec_f64(mats32, f.cpu().numpy(), v.cpu().numpy())
    assert _close(tiled.cpu().numpy(), want)


def test_hub_rows_are_split_and_still_exact(env):
    """Power-law Phi: hub columns give Phi^T rows with thousands of entries; they are cut into chunks
    multiplied by separate groups and summed in order.  Same numbers as float64, with and without the split."""
    eng, torch, o = env["eng"], env["torch"], env["o"]
    lap = o.normalized_laplacian_sparse(powerlaw_graph(5000, 40000, 4))
    g = eng.DeviceGraph.from_scipy(lap)
    phi = eng.build_phi_blocks(g, eng.WalkConfig(40, 0.1, 3, seed=9))
    phi.build_long_rows()
    assert phi._long[1] is not None and phi._long[1]["n_long"] > 0, "test graph should have hub columns"
    n_long, n_chunks = phi._long[1]["n_long"], phi._long[1]["n_chunks"]
    assert n_chunks > n_long
    rng = np.random.default_rng(2)
    f = rng.standard_normal(3).astype(np.float32)
    mats32 = phi.to_scipy_steps()
    for t in (16, 3):
        v = rng.standard_normal((phi.n_rows, t)).astype(np.float32)
        want = o.phi_matvec_f64(mats32, f, v)
        got = phi.matvec(torch.tensor(f), torch.tensor(v).cuda()).cpu().numpy()
        assert _close(got, want)
        got_m = phi.plan(torch.tensor(f), t)(torch.tensor(v).cuda()).cpu().numpy()       # merged Phi_f, own chunks
        assert _close(got_m, want, rtol=5e-5)
    saved, phi._long, phi._long_c = phi._long, [None, None], {}
    unsplit = phi.matvec(torch.tensor(f), torch.tensor(v).cuda()).cpu().numpy()
    phi._long, phi._long_c = saved, {}
    assert _close(unsplit, want) and _close(unsplit, got, rtol=1e-5)


def test_matvec_plan_equals_matvec(env, case):
    """The CG fast paths (one C call per product; per-length blocks or Phi_f merged on the union
    pattern) give the same numbers as the checked path."""
    torch = env["torch"]
    phi = case["phi"]
    gen = torch.Generator(device="cuda").manual_seed(3)
    f = torch.randn(phi.n_steps, device="cuda", generator=gen)
    for t, x1, x2 in [(16, None, None), (5, None, None), (16, torch.arange(0, 600, 3), torch.arange(100, 700, 2))]:
        n2 = phi.n_rows if x2 is None else x2.numel()
        v = torch.randn(n2, t, device="cuda", generator=gen)
        want = phi.matvec(f, v, x1=x1, x2=x2).clone()
        scale = float(want.abs().max())
        plan = phi.plan(f, t, x1=x1, x2=x2, merged=False)
        assert torch.equal(plan(v), want)
        plan.set_modulator(2 * f)
        assert torch.allclose(plan(v), 4 * want, rtol=1e-5, atol=1e-5 * scale)
        mplan = phi.plan(f, t, x1=x1, x2=x2)            # merged (default)
        assert torch.allclose(mplan(v), want, rtol=1e-5, atol=2e-6 * scale)
        mplan.set_modulator(-0.5 * f)
        assert torch.allclose(mplan(v), 0.25 * want, rtol=1e-5, atol=2e-6 * scale)


def test_union_rows_match_scipy_union(env, case):
    """The union pattern / merged values against scipy: Phi_f = sum_l f_l M_l (float32 values)."""
    torch = env["torch"]
    phi = case["phi"]
    f = np.array([0.7, -1.3, 0.4, 2.0], dtype=np.float32)
    mats32 = [m.astype(np.float32) for m in case["mats"]]
    pattern = sum(abs(m).sign() for m in mats32).tocsr()
    assert phi.nnz_union == pattern.nnz
    merged = phi.merged(torch.tensor(f))
    assert merged.n_steps == 1 and merged.nnz == pattern.nnz
    want = sum(float(fl) * m.astype(np.float64) for fl, m in zip(f, mats32)).tocsr()
    want.sort_indices()
    got = merged.to_scipy_steps()[0]
    assert np.array_equal(got.indptr, pattern.indptr) and np.array_equal(got.indices, pattern.indices)
    dense_w = np.zeros(pattern.nnz)
    wd = want.todok()
    rows = np.repeat(np.arange(pattern.shape[0]), np.diff(pattern.indptr))
    ref = np.array(want[rows, pattern.indices]).ravel()
    assert np.allclose(got.data, ref, rtol=1e-5, atol=1e-6 * np.abs(ref).max())
    # the transposed side holds the same matrix
    tptr = merged.tblk_ptr.cpu().numpy()
    tent = merged.tentries.cpu().numpy()
    mt = sp.csr_matrix((tent[:, 1].copy().view(np.float32), tent[:, 0], tptr), shape=(phi.n_cols, phi.n_rows))
    assert abs(mt - got.T).max() == 0


def test_small_training_set_takes_the_scatter_half(env, case):
    """n2 << n_rows (a BO training set): the first half scatters from the selected rows (atomics)."""
    torch, o = env["torch"], env["o"]
    phi = case["phi"]
    rng = np.random.default_rng(21)
    f = rng.standard_normal(phi.n_steps).astype(np.float32)
    mats32 = [m.astype(np.float32) for m in case["mats"]]
    for n2, t in [(30, 1), (30, 16), (7, 5), (1, 3), (40, 20)]:
        assert n2 * 16 < phi.n_rows
        x2 = rng.integers(0, phi.n_rows, size=n2)            # repeats allowed
        x1 = rng.permutation(phi.n_rows)[:333]
        v = rng.standard_normal((n2, t)).astype(np.float32)
        want = o.phi_matvec_f64(mats32, f, v, x1=x1, x2=x2)
        ft, vt = torch.tensor(f), torch.tensor(v).cuda()
        got = phi.matvec(ft, vt, x1=torch.tensor(x1).cuda(), x2=torch.tensor(x2).cuda()).cpu().numpy()
        assert _close(got, want, rtol=5e-5)
        got_p = phi.plan(ft, t, x1=torch.tensor(x1).cuda(), x2=torch.tensor(x2).cuda())(vt).cpu().numpy()
        assert _close(got_p, want, rtol=5e-5)
        # square K[x, x] as in predict(): CG operator
        want_sq = o.phi_matvec_f64(mats32, f, v, x1=x2, x2=x2)
        got_sq = phi.plan(ft, t, x1=torch.tensor(x2).cuda(), x2=torch.tensor(x2).cuda(), merged=False)(vt)
        assert _close(got_sq.cpu().numpy(), want_sq, rtol=5e-5)


def test_apply_and_apply_t_are_the_two_halves(env, case):
    torch = env["torch"]
    phi = case["phi"]
    rng = np.random.default_rng(9)
    f = rng.standard_normal(phi.n_steps).astype(np.float32)
    v = rng.standard_normal((phi.n_rows, 8)).astype(np.float32)
    mats32 = [m.astype(np.float32).astype(np.float64) for m in case["mats"]]
    dense = sum(float(fl) * m for fl, m in zip(f, mats32)).toarray()
    u = phi.apply_t(torch.tensor(f), torch.tensor(v).cuda())
    assert _close(u.cpu().numpy(), dense.T @ v)
    w = phi.apply(torch.tensor(f), u)
    assert _close(w.cpu().numpy(), dense @ (dense.T @ v))


def test_matvec_is_linear_and_symmetric_psd(env, case):
    """Size-independent properties: linearity in V and in f (bilinear), <v, K v> >= 0."""
    torch = env["torch"]
    phi = case["phi"]
    gen = torch.Generator(device="cuda").manual_seed(0)
    f = torch.randn(phi.n_steps, device="cuda", generator=gen)
    a = torch.randn(phi.n_rows, 16, device="cuda", generator=gen)
    b = torch.randn(phi.n_rows, 16, device="cuda", generator=gen)
    ka, kb, kab = phi.matvec(f, a).clone(), phi.matvec(f, b).clone(), phi.matvec(f, 2 * a - 3 * b).clone()
    assert torch.allclose(kab, 2 * ka - 3 * kb, rtol=1e-4, atol=1e-4 * float(kab.abs().max()))
    assert float((a * ka).sum()) >= 0
    assert torch.allclose((a * kb).sum(), (b * ka).sum(), rtol=1e-3)


def test_fgrad_matches_float64_and_finite_differences(env, case):
    torch, o = env["torch"], env["o"]
    phi = case["phi"]
    rng = np.random.default_rng(4)
    f = rng.standard_normal(phi.n_steps).astype(np.float32)
    x1 = rng.permutation(phi.n_rows)[:250]
    x2 = rng.permutation(phi.n_rows)[:310]
    for t in (1, 16, 5):
        left = rng.standard_normal((x1.size, t)).astype(np.float32)
        right = rng.standard_normal((x2.size, t)).astype(np.float32)
        mats32 = [m.astype(np.float32) for m in case["mats"]]
        want = o.phi_fgrad_f64(mats32, f, left, right, x1=x1, x2=x2)
        got = phi.fgrad(torch.tensor(f), torch.tensor(left).cuda(), torch.tensor(right).cuda(),
                        x1=torch.tensor(x1).cuda(), x2=torch.tensor(x2).cuda()).cpu().numpy()
        assert _close(got, want, rtol=1e-4), (t, got, want)
    # the float64 formula itself against central differences
    eps = 1e-6
    left64, right64 = left.astype(np.float64), right.astype(np.float64)
    for l in range(phi.n_steps):
        fp, fm = f.astype(np.float64).copy(), f.astype(np.float64).copy()
        fp[l] += eps
        fm[l] -= eps
        num = (np.sum(left64 * o.phi_matvec_f64(mats32, fp, right64, x1, x2))
               - np.sum(left64 * o.phi_matvec_f64(mats32, fm, right64, x1, x2))) / (2 * eps)
        assert abs(num - want[l]) <= 1e-5 * max(1.0, abs(want[l]))


def test_sharded_rows_sum_to_the_full_product(env, case):
    """Multi-GPU decomposition on one device: row shards' partial U add up, outputs concatenate."""
    eng, torch = env["eng"], env["torch"]
    phi, g, cfg = case["phi"], case["g"], case["cfg"]
    rng = np.random.default_rng(12)
    f = torch.tensor(rng.standard_normal(phi.n_steps).astype(np.float32))
    v = torch.tensor(rng.standard_normal((phi.n_rows, 16)).astype(np.float32)).cuda()
    want = phi.matvec(f, v).clone()
    bounds = [0, 190, 191, 500, phi.n_rows]
    shards = [eng.build_phi_blocks(g, cfg, lo, hi) for lo, hi in zip(bounds[:-1], bounds[1:])]
    u = sum(s.apply_t(f, v[lo:hi]).clone() for s, lo, hi in zip(shards, bounds[:-1], bounds[1:]))
    got = torch.cat([s.apply(f, u) for s in shards])
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))
    # a narrow shard of a banded Phi touches few columns: the non-empty-column list path
    from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian
    lapg = get_normalized_laplacian(grid_graph(60, 50))
    gg = eng.DeviceGraph.from_scipy(lapg)
    cfg2 = eng.WalkConfig(20, 0.1, 3, seed=3)
    full = eng.build_phi_blocks(gg, cfg2)
    part = eng.build_phi_blocks(gg, cfg2, 1000, 1400)
    part.build_long_rows()
    assert part._tcols is not None and part._tcols.numel() < 0.5 * part.n_cols
    vv = torch.tensor(rng.standard_normal((400, 8)).astype(np.float32)).cuda()
    f3 = torch.tensor(rng.standard_normal(3).astype(np.float32))
    vfull_rows = torch.zeros(3000, 8, device="cuda")
    vfull_rows[1000:1400] = vv
    u_part = part.apply_t(f3, vv)
    others = [eng.build_phi_blocks(gg, cfg2, 0, 1000), eng.build_phi_blocks(gg, cfg2, 1400, 3000)]
    u_ref = full.apply_t(f3, vfull_rows)
    assert torch.allclose(u_part, u_ref, rtol=1e-4, atol=1e-5 * float(u_ref.abs().max()))
    # global ids routed to the owning shard only
    x = torch.tensor([3, 190, 450, 699, 200]).cuda()
    parts = torch.zeros(5, 16, device="cuda")
    for s in shards:
        s.apply(f, u, rows=x, out=parts)
    assert torch.allclose(parts, want[x], rtol=1e-4, atol=1e-5 * float(want.abs().max()))


def test_shard_reach_bounds_the_columns_row_shards_share(env, case):
    """grf_shard_reach == (L-1)-hop reachability from every row shard (scipy), and the columns it
    marks as shared contain every column that two shards' Phi blocks actually touch."""
    eng, torch = env["eng"], env["torch"]
    g, cfg, lap = case["g"], case["cfg"], case["lap"]
    n = g.n_nodes
    bounds = [0, 150, 151, 420, n]
    pattern = (abs(lap) > 0).astype(np.int64).tocsr()          # u -> v: v is a neighbour in the walk graph
    L = cfg.max_walk_length
    reach = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        ind = np.zeros(n, dtype=np.int64)
        ind[lo:hi] = 1
        cur = ind.copy()
        for _ in range(L - 1):
            cur = ((pattern.T @ cur) + cur > 0).astype(np.int64)
        reach.append(cur > 0)
    want_shared = np.flatnonzero(np.sum(reach, axis=0) >= 2)
    got = g.shared_columns(bounds, L, max_fraction=2.0)
    assert np.array_equal(got.cpu().numpy(), want_shared)
    assert g.shared_columns(bounds, L, max_fraction=2.0) is got       # cached per (sharding, L)
    # zero hops: every node belongs to exactly one shard
    assert eng.DeviceGraph.from_scipy(lap).shared_columns(bounds, 1, max_fraction=2.0).numel() == 0
    touched = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        mats = eng.build_phi_blocks(g, cfg, lo, hi).to_scipy_steps()
        cols = np.zeros(n, dtype=bool)
        for m in mats:
            cols[m.indices] = True
        touched.append(cols)
        assert not np.any(cols & ~reach[len(touched) - 1])
    actually_shared = np.flatnonzero(np.sum(touched, axis=0) >= 2)
    assert np.isin(actually_shared, want_shared).all()
    # "all" once the shared columns pass the fraction
    assert eng.DeviceGraph.from_scipy(lap).shared_columns(bounds, L, max_fraction=0.0) == "all" or want_shared.size == 0


@pytest.mark.parametrize("kind", ["grid", "powerlaw", "isolated", "ring_bigW", "no_edges"])
def test_every_path_on_small_and_degenerate_graphs(env, kind):
    """Device Laplacian -> walker (chunked staging) -> step matrices / Phi blocks -> every matvec variant
    (per-length, merged plan, row subsets, tiled, fused CG) on small graphs incl. an edgeless one; all
    against float64."""
    eng, torch, o = env["eng"], env["torch"], env["o"]
    from grf_b200.cg import linear_cg_fused
    from gpu_util import ring_graph

    adj, W, L = {"grid": (grid_graph(23, 17), 20, 4), "powerlaw": (powerlaw_graph(1500, 12000, 1), 40, 3),
                 "isolated": (random_graph(200, 90, 2, weighted=True), 7, 5), "ring_bigW": (ring_graph(333), 300, 3),
                 "no_edges": (sp.csr_matrix((40, 40)), 5, 3)}[kind]
    g = eng.DeviceGraph.laplacian_of(adj)
    cfg = eng.WalkConfig(W, 0.1, L, seed=3)
    mats = eng.build_step_matrices(g, cfg, max_stage_bytes=200_000).to_scipy()
    want_mats = env["c"].step_matrices(o.normalized_laplacian_sparse(adj), W, 0.1, L, seed=3)
    for a, b in zip(mats, want_mats):
        assert (abs(a - b)).max() == 0 if a.nnz else b.nnz == 0
    phi = eng.build_phi_blocks(g, cfg)
    n = phi.n_rows
    rng = np.random.default_rng(1)
    f = rng.standard_normal(L).astype(np.float32)
    mats32 = [m.astype(np.float32) for m in mats]
    ft = torch.tensor(f)
    for t in (1, 3, 16, 17, 40):
        v = rng.standard_normal((n, t)).astype(np.float32)
        want = o.phi_matvec_f64(mats32, f, v)
        vt = torch.tensor(v).cuda()
        assert _close(phi.matvec(ft, vt).cpu().numpy(), want, rtol=5e-5)
        assert _close(phi.plan(ft, t)(vt).cpu().numpy(), want, rtol=5e-5)
        x = rng.permutation(n)[: max(1, n // 3)]
        xt = torch.tensor(x).cuda()
        want_s = o.phi_matvec_f64(mats32, f, v[: x.size], x1=x, x2=x)
        assert _close(phi.plan(ft, t, x1=xt, x2=xt)(vt[: x.size].contiguous()).cpu().numpy(), want_s, rtol=5e-5)
    phi.use_tiles = True
    phi.build_windows()
    v = rng.standard_normal((n, 16)).astype(np.float32)
    assert _close(phi.matvec(ft, torch.tensor(v).cuda()).cpu().numpy(), o.phi_matvec_f64(mats32, f, v), rtol=5e-5)
    x = np.sort(rng.permutation(n)[: max(2, n // 2)])
    plan = phi.plan(ft, 8, x1=torch.tensor(x).cuda(), x2=torch.tensor(x).cuda())
    b = rng.standard_normal((x.size, 8)).astype(np.float32)
    sol = linear_cg_fused(plan, torch.tensor(b).cuda(), 0.5, tolerance=1e-6, eps=1e-30, max_iter=400)
    dense = sum(float(fl) * m.astype(np.float64) for fl, m in zip(f, mats32)).toarray()[x]
    want = np.linalg.solve(dense @ dense.T + 0.5 * np.eye(x.size), b.astype(np.float64))
    assert np.abs(sol.cpu().numpy() - want).max() <= 5e-4 * np.abs(want).max()


def test_large_grid_roundtrip_properties(env):
    """BASELINE config-2 shape (316x316 grid, W=100, L=5): M_0 = I, transpose multiset, <a, K b> = <K a, b>."""
    eng, torch = env["eng"], env["torch"]
    from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian

    lap = get_normalized_laplacian(grid_graph(316, 316))
    g = eng.DeviceGraph.from_scipy(lap)
    phi = eng.build_phi_blocks(g, eng.WalkConfig(100, 0.1, 5, seed=42))
    n = phi.n_rows
    assert n == 99856
    ptr = phi.blk_ptr.view(-1)[:-1].view(n, 5)
    assert bool((ptr[:, 1] - ptr[:, 0] == 1).all())                      # one entry per row at length 0
    first = phi.entries[ptr[:, 0].long()]
    assert bool((first[:, 0] == torch.arange(n, device="cuda", dtype=torch.int32)).all())
    assert bool((first[:, 1].view(torch.float32) == 1.0).all())          # M_0 = I exactly
    visits = int(phi.visits)
    expect = n * 100 * sum(0.9 ** k for k in range(5))
    assert abs(visits - expect) / expect < 5e-3
    assert int(phi.tblk_ptr[-1]) == phi.nnz
    gen = torch.Generator(device="cuda").manual_seed(1)
    f = torch.randn(5, device="cuda", generator=gen)
    a = torch.randn(n, 16, device="cuda", generator=gen)
    b = torch.randn(n, 16, device="cuda", generator=gen)
    ka, kb = phi.matvec(f, a).clone(), phi.matvec(f, b).clone()
    lhs, rhs = float((b * ka).sum()), float((a * kb).sum())
    assert abs(lhs - rhs) <= 1e-3 * max(abs(lhs), abs(rhs), 1.0)


def test_row_blocked_transpose_multiplies_like_the_single_block(env, case):
    """A shard beyond ~2^19 rows keeps Phi^T as one block per row range (cache blocking of V; bounded sort
    workspace).  Forced here with tiny blocks: same Phi, the blocks' segments are the single block's segments
    cut by row range, and every product agrees with the single-block result to fp32 summation-order round-off."""
    eng, torch = env["eng"], env["torch"]
    g, cfg, one = case["g"], case["cfg"], case["phi"]
    many = eng.build_phi_blocks(g, cfg, block_rows=150)                        # walked and transposed block by block
    split = eng.build_phi_blocks(g, cfg, transpose=False).build_transpose(block_rows=200)   # one Phi, split afterwards
    assert len(many.tblocks) == 5 and len(split.tblocks) == 4
    L, n, mask = one.n_steps, one.n_cols, (1 << 27) - 1
    tptr1 = one.tblk_ptr.cpu().numpy().astype(np.int64)
    tent1 = one.tentries.cpu().numpy()
    for phi in (many, split):
        assert torch.equal(phi.blk_ptr, one.blk_ptr) and torch.equal(phi.entries, one.entries)
        assert sum(tb.n_rows for tb in phi.tblocks) == one.n_rows and sum(tb.nnz for tb in phi.tblocks) == one.nnz
        got = [[] for _ in range(n * L)]
        for tb in phi.tblocks:
            tp = tb.tblk_ptr.cpu().numpy().astype(np.int64)
            te = tb.tentries.cpu().numpy()
            assert tp[-1] == tb.nnz
            seg_of = np.repeat(np.arange(n * L), np.diff(tp))
            assert np.all((te[:, 0] >> 27) == seg_of % L)
            rows = (te[:, 0] & mask) + tb.r0
            assert rows.min(initial=tb.r0) >= tb.r0 and rows.max(initial=tb.r0) < tb.r0 + tb.n_rows
            order = np.lexsort((rows, seg_of))
            assert np.array_equal(order, np.arange(order.size))               # sorted by (segment, row)
            for k in np.flatnonzero(np.diff(tp)):
                got[k].append(np.stack([rows[tp[k]:tp[k + 1]], te[tp[k]:tp[k + 1], 1]], axis=1))
        for k in range(n * L):
            want = np.stack([tent1[tptr1[k]:tptr1[k + 1], 0] & mask, tent1[tptr1[k]:tptr1[k + 1], 1]], axis=1)
            have = np.concatenate(got[k]) if got[k] else np.zeros((0, 2), dtype=want.dtype)
            assert np.array_equal(have, want), k
    rng = np.random.default_rng(0)
    f = torch.tensor(rng.standard_normal(L).astype(np.float32)).cuda()
    x1 = torch.tensor(rng.integers(0, one.n_rows, 90)).cuda()
    x2 = torch.tensor(np.r_[rng.permutation(one.n_rows)[:300], [5, 5, 7]]).cuda()          # repeated ids too
    for t in (1, 4, 16, 17):
        v = torch.tensor(rng.standard_normal((one.n_rows, t)).astype(np.float32)).cuda()
        v2 = torch.tensor(rng.standard_normal((x2.numel(), t)).astype(np.float32)).cuda()
        want_u, want = one.apply_t(f, v).clone(), one.matvec(f, v).clone()
        want_sub = one.matvec(f, v2, x1=x1, x2=x2).clone()
        want_plan = one.plan(f, t, merged=False)(v).clone()
        for phi in (many, split):
            tol = 4e-6 * float(want_u.abs().max())
            assert float((phi.apply_t(f, v) - want_u).abs().max()) <= tol
            assert float((phi.matvec(f, v) - want).abs().max()) <= 4e-6 * float(want.abs().max())
            assert float((phi.matvec(f, v2, x1=x1, x2=x2) - want_sub).abs().max()) <= 4e-6 * float(want_sub.abs().max())
            plan = phi.plan(f, t)                     # merged=True falls back to the per-length blocks here
            assert not plan.merged
            assert float((plan(v) - want_plan).abs().max()) <= 4e-6 * float(want_plan.abs().max())
            sub = phi.plan(f, t, x1=x1, x2=x2[:300])
            assert float((sub(v2[:300]) - one.matvec(f, v2[:300], x1=x1, x2=x2[:300])).abs().max()) <= \
                4e-6 * float(want_sub.abs().max())
    left = torch.tensor(rng.standard_normal((one.n_rows, 4)).astype(np.float32)).cuda()
    right = torch.tensor(rng.standard_normal((one.n_rows, 4)).astype(np.float32)).cuda()
    gw = one.fgrad(f, left, right)
    assert float((many.fgrad(f, left, right) - gw).abs().max()) <= 1e-4 * float(gw.abs().max())


def test_streaming_gathers_are_bit_identical(env, case):
    """L1::no_allocate gathers (GrfPhi flag +32; picked per Phi by MatvecPlan._tune_gathers) change where a
    gathered row is cached, not what is computed."""
    eng, torch = env["eng"], env["torch"]
    phi = eng.build_phi_blocks(case["g"], case["cfg"])
    rng = np.random.default_rng(1)
    f = torch.tensor(rng.standard_normal(phi.n_steps).astype(np.float32)).cuda()
    for t in (3, 4, 16, 40):
        v = torch.tensor(rng.standard_normal((phi.n_rows, t)).astype(np.float32)).cuda()
        phi.stream_gather = (False, False)
        want = phi.matvec(f, v).clone()
        for mode in ((True, True), (True, False), (False, True)):
            phi.stream_gather = mode
            assert torch.equal(phi.matvec(f, v), want), (t, mode)
            plan = phi.plan(f, t, merged=False)
            assert torch.equal(plan(v), want), (t, mode)
    plan = phi.plan(f, 16, merged=False)
    assert plan._tune_gathers() in ((False, False), (False, True), (True, False), (True, True))


def test_torch_csr_with_unsorted_or_repeated_columns_is_coalesced():
    """torch accepts CSR tensors whose rows are unsorted or repeat a column (its SpMM adds the repeats); the Phi
    blocks built from one must multiply like torch does, merged layout included."""
    import torch
    from grf_b200 import engine

    crow = torch.tensor([0, 3, 3, 6, 8])
    col = torch.tensor([2, 0, 2, 3, 1, 0, 1, 1])          # row 0: unsorted + repeated 2; row 3: repeated 1
    val = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0])
    m1 = torch.sparse_csr_tensor(crow, col, val, (4, 4), dtype=torch.float32).cuda()
    m0 = torch.eye(4).to_sparse_csr().cuda()
    dense = [m0.to_dense().double(), torch.zeros(4, 4, dtype=torch.float64).cuda()]
    for r in range(4):
        for k in range(int(crow[r]), int(crow[r + 1])):
            dense[1][r, int(col[k])] += float(val[k])
    phi = engine.phi_blocks_from_torch_csr([m0, m1])
    f = torch.tensor([0.5, -1.5]).cuda()
    pf = f[0].double() * dense[0] + f[1].double() * dense[1]
    v = torch.randn(4, 3).cuda()
    want = (pf @ (pf.T @ v.double())).float()
    for merged in (False, True):
        got = phi.plan(f, 3, merged=merged)(v)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), merged


def test_union_layout_with_rows_longer_than_one_chunk(env):
    """The union (merged) layout splits rows beyond engine.UNION_CHUNK entries into chunk tasks: a star's hub column
    holds ~n entries per length in Phi^T, and the hub's own row of Phi holds one per leaf.  Phi_f on the union
    pattern must equal sum_l f_l M_l entry for entry, and multiply like the per-length blocks."""
    torch, eng = env["torch"], env["eng"]
    n = 9000
    leaves = np.arange(1, n)
    adj = sp.csr_matrix((np.ones(2 * (n - 1)), (np.r_[np.zeros(n - 1, dtype=np.int64), leaves],
                                                np.r_[leaves, np.zeros(n - 1, dtype=np.int64)])), shape=(n, n))
    extra = random_graph(n, 3 * n, 17)
    lap = env["o"].normalized_laplacian_sparse(((adj + extra) > 0).astype(float).tocsr())
    g = eng.DeviceGraph.from_scipy(lap)
    phi = eng.build_phi_blocks(g, eng.WalkConfig(12, 0.1, 4, seed=9))
    seg = np.diff(phi.tblk_ptr.cpu().numpy().astype(np.int64)[::4])
    assert seg.max() > 3 * eng.UNION_CHUNK, "the hub column should span several chunks"
    f = torch.tensor([0.7, -1.3, 0.4, 2.0], device="cuda")
    merged = phi.merged(f)
    mats = phi.to_scipy_steps()                                     # float32 values
    want = sum(float(fl) * m.astype(np.float64) for fl, m in zip(f.cpu().numpy(), mats)).tocsr()
    want.sort_indices()
    got = merged.to_scipy_steps()[0].astype(np.float64)
    got.sort_indices()
    assert phi.nnz_union == got.nnz
    # the union pattern keeps a column even when its f-weighted sum cancels; compare as dense-free difference
    diff = (got - want)
    assert abs(diff).max() <= 1e-6 * abs(want).max()
    got_t = sp.csr_matrix((merged.tentries.cpu().numpy()[:, 1].copy().view(np.float32).astype(np.float64),
                           merged.tentries.cpu().numpy()[:, 0] & ((1 << 27) - 1),
                           merged.tblk_ptr.cpu().numpy().astype(np.int64)), shape=(n, n))
    assert abs(got_t - want.T).max() <= 1e-6 * abs(want).max()
    v = torch.randn(n, 5, device="cuda")
    a = phi.plan(f, 5, merged=False)(v)
    b = phi.plan(f, 5, merged=True)(v)
    assert float((a - b).abs().max()) <= RTOL * float(a.abs().max())


@pytest.mark.parametrize("shape", [(37, 29), (64, 48)])
def test_line_pair_layout_multiplies_like_the_merged_entries(env, shape, monkeypatch):
    """csrc/grf_pairs.cu: on a lattice most columns of a row have their line partner (col ^ 1); the t = 16 merged
    product then runs on pair entries.  Same result as the plain merged product (same weights, one more
    rounding order), as the per-length product and as float64; odd row counts, a row subset x1, modulator updates."""
    torch, eng = env["torch"], env["eng"]
    lap = env["o"].normalized_laplacian_sparse(grid_graph(*shape))
    n = lap.shape[0]
    g = eng.DeviceGraph.from_scipy(lap)
    phi = eng.build_phi_blocks(g, eng.WalkConfig(30, 0.1, 4, seed=3))
    assert phi.pair_ratio < 0.75, phi.pair_ratio
    f = torch.tensor([1.0, -0.6, 0.3, 0.2], device="cuda")
    v = torch.randn(n, 16, device="cuda")
    plain = phi.plan(f, 16, merged=True)
    assert plain._pair is None or eng.PAIR_LAYOUT            # opt-in (GRF_PAIR_LAYOUT=1)
    want = plain(v)
    monkeypatch.setattr(eng, "PAIR_LAYOUT", True)
    plan = phi.plan(f, 16, merged=True)
    assert plan._pair is not None
    got = plan(v)
    scale = float(want.abs().max())
    assert float((got - want).abs().max()) <= RTOL * scale
    mats = phi.to_scipy_steps()
    pf = sum(float(fl) * m.astype(np.float64) for fl, m in zip(f.cpu().numpy(), mats))
    ref = pf @ (pf.T @ v.cpu().numpy().astype(np.float64))
    assert _close(got.cpu().numpy(), ref)
    # a row subset on the output side, and a new modulator through the same plan
    x1 = torch.arange(1, n, 3, device="cuda")
    sub = phi.plan(f, 16, x1=x1, merged=True)
    assert sub._pair is not None
    assert _close(sub(v).cpu().numpy(), ref[x1.cpu().numpy()])
    f2 = torch.tensor([0.5, 0.25, -1.0, 0.7], device="cuda")
    plan.set_modulator(f2)
    pf2 = sum(float(fl) * m.astype(np.float64) for fl, m in zip(f2.cpu().numpy(), mats))
    assert _close(plan(v).cpu().numpy(), pf2 @ (pf2.T @ v.cpu().numpy().astype(np.float64)))
    # operands the pair kernel cannot take (a misaligned view of V) fall back to the merged entries
    big = torch.randn(n + 1, 16, device="cuda")
    assert _close(plan(big[1:]).cpu().numpy(), pf2 @ (pf2.T @ big[1:].cpu().numpy().astype(np.float64)))


def test_line_pair_layout_is_not_chosen_on_power_law_graphs(env, monkeypatch):
    torch, eng = env["torch"], env["eng"]
    monkeypatch.setattr(eng, "PAIR_LAYOUT", True)
    lap = env["o"].normalized_laplacian_sparse(powerlaw_graph(3000, 20000, 4))
    phi = eng.build_phi_blocks(eng.DeviceGraph.from_scipy(lap), eng.WalkConfig(20, 0.1, 3, seed=1))
    plan = phi.plan(torch.ones(3, device="cuda"), 16, merged=True)
    assert plan._pair is None


def test_threaded_pinned_upload_equals_the_plain_copy(env, monkeypatch):
    """engine._upload: large host arrays leave pageable memory through pinned slots filled by several host threads
    (chunks in flight on per-thread streams); forced here on a small graph, odd sizes included."""
    torch, eng = env["torch"], env["eng"]
    adj = random_graph(5000, 40011, 3, weighted=True)
    plain = eng.DeviceGraph.from_scipy(adj)
    monkeypatch.setattr(eng, "_UPLOAD_MIN", 1 << 10)
    monkeypatch.setattr(eng, "_UPLOAD_CHUNK", 7 << 10)          # many chunks per array, the last one partial
    eng._upload_state.clear()
    try:
        threaded = eng.DeviceGraph.from_scipy(adj)
        for a, b in ((plain.row_ptr, threaded.row_ptr), (plain.col_idx, threaded.col_idx), (plain.val, threaded.val)):
            assert torch.equal(a, b)
    finally:
        eng._upload_state.clear()                                  # slots of the test's chunk size must not survive


def test_pilot_balanced_shards_even_out_the_entries(env):
    """sharding.pilot_row_cost / balanced_bounds: contiguous, covering, and on a power-law graph closer to equal
    entries per shard than the degree-only estimate."""
    torch, eng = env["torch"], env["eng"]
    from grf_b200 import sharding, synth

    g, _ = synth.rmat_walk_graph(15, 300_000, seed=2, device=torch.device("cuda", 0))
    cfg = eng.WalkConfig(60, 0.1, 4, seed=3)
    full = eng.build_phi_blocks(g, cfg, transpose=False)
    ptr = full.blk_ptr.cpu().numpy().astype(np.int64)[::4]

    def spread(bounds):
        assert bounds[0] == 0 and bounds[-1] == g.n_nodes and all(a <= b for a, b in zip(bounds, bounds[1:]))
        nnz = np.array([ptr[b] - ptr[a] for a, b in zip(bounds, bounds[1:])], dtype=float)
        return nnz.max() / nnz.mean()

    by_degree = spread(sharding.balanced_bounds(g, 8))
    by_pilot = spread(sharding.balanced_bounds(g, 8, row_cost=sharding.pilot_row_cost(g, cfg, pilot_walks=15)))
    assert by_pilot < by_degree and by_pilot < 1.15, (by_degree, by_pilot)


def test_config4_full_size_properties(env):
    """BASELINE config 4 at full size (R-MAT 2^22 nodes, >= 70 M edges, W = 100, L = 5, t = 16), through
    size-independent properties: M_0 = I, Phi^T is a permutation of Phi (sums of a hash of (row, column, length,
    value) over both layouts, and every segment row-sorted on a sample), the product is linear and symmetric
    (<a, K b> = <K a, b>), the union layout multiplies like the per-length blocks, and a row shard rebuilt on its
    own reproduces its rows bit for bit (counter-based draws)."""
    torch, eng = env["torch"], env["eng"]
    from grf_b200 import synth

    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~50 GB of device memory")
    dev = torch.device("cuda", 0)
    g, stats = synth.rmat_walk_graph(22, 70_000_000, seed=0, device=dev)
    assert stats["n_nodes"] == 1 << 22 and stats["undirected_edges"] >= 70_000_000
    cfg = eng.WalkConfig(100, 0.1, 5, seed=42)
    phi = eng.build_phi_blocks(g, cfg)
    n, L, mask = phi.n_rows, 5, (1 << 27) - 1
    ptr = phi.blk_ptr.view(-1)[:-1].view(n, L)
    assert bool((ptr[:, 1] - ptr[:, 0] == 1).all())
    first = phi.entries[ptr[:, 0].long()]
    assert bool((first[:, 0] == torch.arange(n, device=dev, dtype=torch.int32)).all())
    assert bool((first[:, 1].view(torch.float32) == 1.0).all())                       # M_0 = I exactly
    assert int(phi.tblk_ptr[-1]) == phi.nnz > 500_000_000

    def digest(ptr_all, ent, rows_are_owner):
        # sum over entries of a 61-bit mix of (row, column, length, value bits); order-independent
        owner = torch.repeat_interleave(torch.arange(ptr_all.numel() - 1, device=dev),
                                        (ptr_all[1:] - ptr_all[:-1]).long(), output_size=ent.shape[0]) // L
        other = (ent[:, 0] & mask).long()
        step = (ent[:, 0] >> 27).long() & 31
        row, col = (owner, other) if rows_are_owner else (other, owner)
        h = (row * 1000003 + col) * 31 + step
        h = (h * 2654435761 + ent[:, 1].long() * 97) & ((1 << 61) - 1)
        return int(h.sum() & ((1 << 62) - 1)), int((h ^ (h >> 17)).sum() & ((1 << 62) - 1))

    assert digest(phi.blk_ptr, phi.entries, True) == digest(phi.tblk_ptr, phi.tentries, False)
    tptr = phi.tblk_ptr
    seg = (tptr[1:] - tptr[:-1])
    for s in torch.topk(seg, 3).indices.tolist() + torch.randint(0, seg.numel(), (200,), device=dev).tolist():
        r = phi.tentries[int(tptr[s]):int(tptr[s + 1]), 0] & mask
        assert bool((r[1:] > r[:-1]).all())
    gen = torch.Generator(device=dev).manual_seed(1)
    f = torch.randn(L, device=dev, generator=gen)
    a = torch.randn(n, 16, device=dev, generator=gen)
    b = torch.randn(n, 16, device=dev, generator=gen)
    plan = phi.plan(f, 16, merged=False)
    ka, kb = plan(a).clone(), plan(b).clone()
    lhs, rhs = float((b.double() * ka.double()).sum()), float((a.double() * kb.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs))
    kab = plan(2.5 * a + b)
    assert float((kab - (2.5 * ka + kb)).abs().max()) <= RTOL * float(kab.abs().max())
    merged = phi.plan(f, 16, merged=True)(a)
    assert float((merged - ka).abs().max()) <= RTOL * float(ka.abs().max())
    del plan, merged, kab, kb
    lo, hi = 1_000_000, 1_050_000
    part = eng.build_phi_blocks(g, cfg, lo, hi, transpose=False)
    b0, b1 = int(phi.blk_ptr[lo * L]), int(phi.blk_ptr[hi * L])
    assert torch.equal(part.entries, phi.entries[b0:b1])
    assert torch.equal(part.blk_ptr, phi.blk_ptr[lo * L:hi * L + 1] - b0)
