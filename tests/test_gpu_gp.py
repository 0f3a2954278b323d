"""-m gpu tests of the GPyTorch-facing drop-ins: GraphPreprocessor, SparseLinearOperator,
SparseGRFKernel / SparseDiffusionKernel (forward, autograd through the modulator), on-device
CG and SparseGraphGP.predict.  gpytorch / linear_operator are not installed here, so the
kernels run on the stand-in protocol of grf_b200.linop / gp_compat (parity for this layer is
against float64 numpy, see oracle/__init__.py)."""

import numpy as np
import pytest
import scipy.sparse as sp

from gpu_util import csr_bits_equal, grid_graph, random_graph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    import torch

    assert torch.cuda.is_available()
    from efficient_graph_gp_sparse.preprocessor import GraphPreprocessor

    adj = random_graph(400, 1300, 21)
    pp = GraphPreprocessor(adj, walks_per_node=30, p_halt=0.1, max_walk_length=4, random_walk_seed=7, use_tqdm=False)
    ops = pp.preprocess_graph()
    return dict(torch=torch, adj=adj, pp=pp, ops=ops)


def _phi64(mats, f):
    return sum(float(fl) * m.astype(np.float32).astype(np.float64) for fl, m in zip(f, mats)).toarray()


def test_preprocessor_matches_sampler_and_keeps_reference_surface(setup, tmp_path):
    from efficient_graph_gp_sparse.preprocessor import GraphPreprocessor
    from efficient_graph_gp_sparse.random_walk_samplers_sparse import SparseRandomWalk
    from efficient_graph_gp_sparse.utils_sparse import SparseLinearOperator, get_normalized_laplacian

    pp, ops = setup["pp"], setup["ops"]
    assert len(ops) == 4 and all(isinstance(o, SparseLinearOperator) for o in ops)
    assert len(pp.step_matrices_scipy) == 4 and pp.step_matrices_torch is ops
    direct = SparseRandomWalk(get_normalized_laplacian(setup["adj"]), seed=7).get_random_walk_matrices(30, 0.1, 4)
    for a, b in zip(pp.step_matrices_scipy, direct):
        assert csr_bits_equal(a, b)
    t = ops[1].sparse_csr_tensor
    assert t.is_sparse_csr and t.dtype == setup["torch"].float32 and t.crow_indices().dtype == setup["torch"].int64
    assert ops.phi_blocks is not None
    # pickle cache: same format as the reference (a list of scipy CSR), round trip
    cache = str(tmp_path / "steps.pkl")
    pp.save_step_matrices(pp.step_matrices_scipy, cache)
    again = GraphPreprocessor(setup["adj"], 30, 0.1, 4, 7, load_from_disk=True, cache_filename=cache)
    for a, b in zip(again.step_matrices_scipy, pp.step_matrices_scipy):
        assert csr_bits_equal(a, b)
    with pytest.raises(FileNotFoundError):
        GraphPreprocessor(setup["adj"], load_from_disk=True, cache_filename=str(tmp_path / "missing.pkl"))
    with pytest.raises(ValueError, match="square"):
        GraphPreprocessor(sp.csr_matrix((3, 4)))
    with pytest.raises(ValueError, match="scipy CSR"):
        GraphPreprocessor.from_scipy_csr(np.eye(3))


def test_sparse_linear_operator_protocol(setup):
    """sparse_lo.py:4-25 -- _matmul, _size, _transpose_nonbatch (self-check of sparse_lo.py:53-60: SpMM == dense)."""
    torch = setup["torch"]
    from efficient_graph_gp_sparse.utils_sparse import SparseLinearOperator

    op = setup["ops"][2]
    dense = setup["pp"].step_matrices_scipy[2].astype(np.float32).astype(np.float64).toarray()
    rhs = torch.randn(400, 5, device="cuda")
    assert tuple(op._size()) == (400, 400)
    got = op._matmul(rhs).cpu().numpy()
    assert np.allclose(got, dense @ rhs.cpu().numpy().astype(np.float64), rtol=1e-5, atol=1e-5)
    got_t = op._transpose_nonbatch()._matmul(rhs).cpu().numpy()
    assert np.allclose(got_t, dense.T @ rhs.cpu().numpy().astype(np.float64), rtol=1e-5, atol=1e-5)
    assert op._transpose_nonbatch()._transpose_nonbatch() is op
    assert torch.allclose(op._transpose_nonbatch().sparse_csr_tensor.to_dense(), op.sparse_csr_tensor.to_dense().T)
    with pytest.raises(ValueError, match="CSR"):
        SparseLinearOperator(torch.eye(3).cuda())


def test_grf_kernel_forward_and_modulator_gradient(setup):
    torch = setup["torch"]
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseGRFKernel

    torch.manual_seed(42)
    kern = SparseGRFKernel(max_walk_length=4, step_matrices_torch=setup["ops"]).cuda()
    assert kern.raw_modulator_vector.shape == (4,) and kern.modulator_vector is kern.raw_modulator_vector
    f = kern.modulator_vector.detach().cpu().numpy()
    phi = _phi64(setup["pp"].step_matrices_scipy, f)
    rng = np.random.default_rng(1)
    x1 = torch.tensor(rng.permutation(400)[:150], dtype=torch.float32)[:, None].cuda()   # float column, as the models pass it
    x2 = torch.tensor(rng.permutation(400)[:90], dtype=torch.float32)[:, None].cuda()
    K = kern(x1, x2)
    assert tuple(K.shape) == (150, 90)
    i1, i2 = x1.long().flatten().cpu().numpy(), x2.long().flatten().cpu().numpy()
    want = phi[i1] @ phi[i2].T
    assert np.allclose(K.to_dense().detach().cpu().numpy(), want, rtol=1e-4, atol=1e-4 * np.abs(want).max())
    # K @ v through the lazy operator, and the gradient of a scalar of it w.r.t. the modulator
    v = torch.randn(90, 8, device="cuda")
    g_out = torch.randn(150, 8, device="cuda")
    loss = (g_out * (K @ v)).sum()
    loss.backward()
    from oracle import grf_oracle as orc

    want_grad = orc.phi_fgrad_f64([m.astype(np.float32) for m in setup["pp"].step_matrices_scipy], f,
                                  g_out.cpu().numpy(), v.cpu().numpy(), x1=i1, x2=i2)
    got_grad = kern.raw_modulator_vector.grad.cpu().numpy()
    assert np.allclose(got_grad, want_grad, rtol=2e-4, atol=2e-4 * np.abs(want_grad).max())
    # full kernel (no index), diag, symmetry / PSD as the reference's tests check for K
    Kfull = kern.forward().to_dense().detach().cpu().numpy()
    assert np.allclose(Kfull, Kfull.T, atol=1e-4) and np.linalg.eigvalsh(Kfull.astype(np.float64)).min() > -1e-3
    d = kern.forward(x1, x1, diag=True).detach().cpu().numpy()
    assert np.allclose(d, np.sum(phi[i1] ** 2, axis=1), rtol=1e-4, atol=1e-4)


def test_diffusion_kernel_modulator_and_gradients(setup):
    torch = setup["torch"]
    from efficient_graph_gp.modulation_functions import diffusion_modulator
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseDiffusionKernel, diffusion_modulator_torch

    kern = SparseDiffusionKernel(max_walk_length=4, step_matrices_torch=setup["ops"]).cuda()
    beta, sigma = float(kern.beta), float(kern.sigma_f)
    assert beta > 0 and sigma > 0
    want = np.array([sigma * diffusion_modulator(l, beta) for l in range(4)])
    assert np.allclose(kern.modulator_vector.detach().cpu().numpy(), want, rtol=1e-5)
    assert np.allclose(diffusion_modulator_torch(torch.arange(4), torch.tensor(2.0)).numpy(),
                       [diffusion_modulator(l, 2.0) for l in range(4)], rtol=1e-6)
    x = torch.arange(0, 400, 3).cuda()
    v = torch.randn(x.numel(), 4, device="cuda")
    loss = (v * (kern(x, x) @ v)).sum()
    loss.backward()
    assert kern.raw_beta.grad is not None and kern.raw_sigma_f.grad is not None
    # d loss / d sigma_f = 2 loss / sigma_f (K is quadratic in sigma_f); chain through softplus
    dsig = 2 * float(loss) / sigma * float(torch.sigmoid(kern.raw_sigma_f))
    assert abs(float(kern.raw_sigma_f.grad) - dsig) <= 2e-3 * abs(dsig)


def test_cg_posterior_mean_within_1e4_of_direct_solve(setup):
    """North star: CG posterior means agree to within 1e-4 relative in fp32 (vs a float64 direct solve)."""
    torch = setup["torch"]
    from efficient_graph_gp_sparse.models import SparseGraphGP
    from grf_b200.gp_compat import GaussianLikelihood

    torch.manual_seed(0)
    rng = np.random.default_rng(3)
    train = rng.permutation(400)[:240]
    test = np.setdiff1d(np.arange(400), train)
    y = np.sin(np.arange(400) / 20.0)[train] + 0.1 * rng.standard_normal(240)
    lik = GaussianLikelihood()
    lik.noise = 0.1
    model = SparseGraphGP(torch.tensor(train, dtype=torch.float32)[:, None].cuda(),
                          torch.tensor(y, dtype=torch.float32).cuda(), lik, setup["ops"], 4).cuda()
    f = model.covar_module.modulator_vector.detach().cpu().numpy()
    phi = _phi64(setup["pp"].step_matrices_scipy, f)
    Ktt = phi[train] @ phi[train].T + float(lik.noise) * np.eye(240)
    want = phi[test] @ phi[train].T @ np.linalg.solve(Ktt, y.astype(np.float32).astype(np.float64))
    got, info = model.posterior_mean(torch.tensor(test).cuda(), cg_tolerance=1e-7, return_info=True)
    rel = np.linalg.norm(got.cpu().numpy() - want) / np.linalg.norm(want)
    assert rel <= 1e-4, (rel, info)
    # pathwise samples: right shape, finite, and centred on the mean
    one, _ = model.predict(torch.tensor(test).cuda(), n_samples=1, cg_tolerance=1e-4, return_info=True)   # Thompson
    assert tuple(one.shape) == (1, test.size) and bool(torch.isfinite(one).all())
    samples, info = model.predict(torch.tensor(test).cuda(), n_samples=200, cg_tolerance=1e-4, return_info=True)
    assert tuple(samples.shape) == (200, test.size) and bool(torch.isfinite(samples).all())
    post_var = np.diag(phi[test] @ phi[test].T - phi[test] @ phi[train].T @ np.linalg.solve(Ktt, phi[train] @ phi[test].T))
    z = (samples.mean(0).cpu().numpy() - want) / np.sqrt(np.maximum(post_var, 1e-12) / 200)
    assert np.mean(np.abs(z) < 4) > 0.97


def test_mll_gradient_matches_dense_float64(setup):
    """-mll/n gradient w.r.t. the modulator and the noise: CG + per-length reductions vs a dense float64
    evaluation.  Probes = sqrt(n) e_j (p = n) make the trace estimator exact, so the comparison is tight."""
    torch = setup["torch"]
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseGRFKernel
    from grf_b200.gp_compat import GaussianLikelihood
    from grf_b200.mll import neg_mll_backward

    torch.manual_seed(1)
    rng = np.random.default_rng(11)
    kern = SparseGRFKernel(4, setup["ops"]).cuda()
    with torch.no_grad():
        kern.raw_modulator_vector.copy_(torch.tensor([1.0, -0.4, 0.2, -0.05]))
    lik = GaussianLikelihood()
    lik.noise = 0.5
    idx = np.sort(rng.permutation(400)[:120])
    y = np.sin(idx / 15.0) + 0.1 * rng.standard_normal(120)
    n = 120
    probes = torch.eye(n) * np.sqrt(n)
    out = neg_mll_backward(kern, lik, torch.tensor(idx), torch.tensor(y), probes=probes, cg_tolerance=1e-7,
                           cg_eps=1e-30)
    # dense float64 reference
    f = kern.modulator_vector.detach().cpu().numpy().astype(np.float64)
    mats = [m.astype(np.float32).astype(np.float64)[idx].toarray() for m in setup["pp"].step_matrices_scipy]
    phi = sum(fl * m for fl, m in zip(f, mats))
    s2 = float(lik.noise)
    Kh = phi @ phi.T + s2 * np.eye(n)
    Kinv = np.linalg.inv(Kh)
    a = Kinv @ y.astype(np.float32).astype(np.float64)
    want_f = []
    for l in range(4):
        dK = mats[l] @ phi.T + phi @ mats[l].T
        want_f.append(-(0.5 * a @ dK @ a - 0.5 * np.trace(Kinv @ dK)) / n)
    want_s2 = -(0.5 * a @ a - 0.5 * np.trace(Kinv)) / n
    got_f = kern.raw_modulator_vector.grad.cpu().numpy()
    assert np.allclose(got_f, want_f, rtol=2e-3, atol=2e-3 * np.abs(want_f).max()), (got_f, want_f)
    # noise gradient reaches raw_noise through the softplus constraint
    raw = float(lik.raw_noise)
    dnoise_draw = 1.0 / (1.0 + np.exp(-raw))
    assert abs(float(lik.raw_noise.grad) - want_s2 * dnoise_draw) <= 2e-3 * abs(want_s2 * dnoise_draw) + 1e-6
    assert abs(out["datafit"] - 0.5 * y.astype(np.float32) @ a) <= 1e-3 * abs(0.5 * y @ a)
    # loss value: with probes sqrt(n) e_j and enough Lanczos steps the quadrature is exact
    from grf_b200.mll import lanczos_logdet
    Kop = kern(torch.tensor(idx).cuda(), torch.tensor(idx).cuda())
    plan = Kop.plan(n)
    ld = lanczos_logdet(lambda v: plan(v.contiguous()) + s2 * v, probes.cuda(), iterations=40)
    sign, want_ld = np.linalg.slogdet(Kh)
    assert abs(ld - want_ld) <= 2e-3 * abs(want_ld), (ld, want_ld)
    want_loss = (0.5 * y.astype(np.float32) @ a + 0.5 * want_ld + 0.5 * n * np.log(2 * np.pi)) / n
    ld1 = lanczos_logdet(lambda v: plan(v.contiguous()) + s2 * v, probes.cuda(), iterations=1)   # the reference's setting
    assert abs(ld1 - np.sum(np.log(np.diag(Kh)))) <= 1e-3 * abs(want_ld)
    assert out["loss"] is not None and np.isfinite(out["loss"])
    # with random probes the estimate is unbiased: average of a few draws lands near the exact value
    kern.raw_modulator_vector.grad = None
    lik.raw_noise.grad = None
    gen = torch.Generator(device="cuda").manual_seed(5)
    acc = np.zeros(4)
    for _ in range(8):
        kern.raw_modulator_vector.grad = None
        neg_mll_backward(kern, lik, torch.tensor(idx), torch.tensor(y), num_probes=64, cg_tolerance=1e-6,
                         generator=gen, cg_eps=1e-30)
        acc += kern.raw_modulator_vector.grad.cpu().numpy() / 8
    assert np.allclose(acc, want_f, rtol=0.15, atol=0.1 * np.abs(want_f).max()), (acc, want_f)


def test_fused_cg_matches_unfused_cg_and_direct_solve(setup):
    """csrc/grf_cg.cu: same iterates as the torch-op CG, same answer as a float64 direct solve."""
    torch = setup["torch"]
    from grf_b200.cg import linear_cg, linear_cg_fused

    blocks = setup["ops"].phi_blocks
    rng = np.random.default_rng(8)
    f = torch.tensor(rng.standard_normal(4).astype(np.float32)).cuda()
    x = torch.tensor(rng.permutation(400)[:230]).cuda()
    sigma2 = 0.3
    phi = _phi64(setup["pp"].step_matrices_scipy, f.cpu().numpy())
    idx = x.cpu().numpy()
    a = phi[idx] @ phi[idx].T + sigma2 * np.eye(230)
    for t in (1, 5, 16, 40):
        b = rng.standard_normal((230, t)).astype(np.float32)
        plan = blocks.plan(f, t, x1=x, x2=x)
        got, info = linear_cg_fused(plan, torch.tensor(b).cuda(), sigma2, tolerance=1e-6, eps=1e-30, check_every=1,
                                    return_info=True)
        ref, info2 = linear_cg(lambda v: plan(v) + sigma2 * v, torch.tensor(b).cuda(), tolerance=1e-6, eps=1e-30,
                               check_every=1, return_info=True)
        want = np.linalg.solve(a, b.astype(np.float64))
        scale = np.abs(want).max()
        assert np.abs(got.cpu().numpy() - want).max() <= 2e-4 * scale, (t, info)
        assert np.abs(got.cpu().numpy() - ref.cpu().numpy()).max() <= 2e-4 * scale
        assert abs(info["iterations"] - info2["iterations"]) <= 3
    with pytest.raises(ValueError, match="square"):
        linear_cg_fused(blocks.plan(f, 4, x1=x, x2=x[:100]), torch.zeros(100, 4).cuda())


def test_device_cg_matches_oracle_cg_semantics(setup):
    torch = setup["torch"]
    from grf_b200.cg import linear_cg
    from oracle import grf_oracle as orc

    rng = np.random.default_rng(5)
    a = rng.standard_normal((60, 60))
    a = a @ a.T + 60 * np.eye(60)
    b = rng.standard_normal((60, 3))
    want = orc.linear_cg(lambda v: a @ v, b, tolerance=1e-8, eps=1e-30)
    at = torch.tensor(a, dtype=torch.float32).cuda()
    got = linear_cg(lambda v: at @ v, torch.tensor(b, dtype=torch.float32).cuda(), tolerance=1e-7, check_every=1, eps=1e-30)
    assert np.allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-5)
    assert np.allclose(want, np.linalg.solve(a, b), rtol=1e-6, atol=1e-8)


def test_diag_is_per_pair_sparse_dots_with_modulator_gradient(setup):
    """diag=True (sparse_grf_kernel.py:55-57) through grf_phi_row_dots: values and d/df against a float64
    dense evaluation, for x1 == x2, for two different index lists and for the un-indexed kernel."""
    torch = setup["torch"]
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseGRFKernel

    torch.manual_seed(3)
    kern = SparseGRFKernel(4, setup["ops"]).cuda()
    f = kern.modulator_vector.detach().cpu().numpy().astype(np.float64)
    mats = [m.astype(np.float32).astype(np.float64).toarray() for m in setup["pp"].step_matrices_scipy]
    phi = sum(fl * m for fl, m in zip(f, mats))
    rng = np.random.default_rng(5)
    i1, i2 = rng.permutation(400)[:120], rng.permutation(400)[:120]
    for a, b in ((i1, i1), (i1, i2)):
        kern.raw_modulator_vector.grad = None
        d = kern.forward(torch.tensor(a).cuda(), torch.tensor(b).cuda(), diag=True)
        want = np.sum(phi[a] * phi[b], axis=1)
        assert tuple(d.shape) == (120,)
        assert np.allclose(d.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
        g = torch.tensor(rng.standard_normal(120).astype(np.float32)).cuda()
        (d * g).sum().backward()
        gw = g.cpu().numpy().astype(np.float64)
        want_grad = [np.sum(gw * (np.sum(mats[l][a] * phi[b], axis=1) + np.sum(phi[a] * mats[l][b], axis=1)))
                     for l in range(4)]
        got_grad = kern.raw_modulator_vector.grad.cpu().numpy()
        assert np.allclose(got_grad, want_grad, rtol=1e-4, atol=1e-4 * np.abs(want_grad).max())
    full = kern.forward(diag=True).detach().cpu().numpy()
    assert np.allclose(full, np.sum(phi * phi, axis=1), rtol=1e-5, atol=1e-5)
    K = kern(torch.tensor(i1).cuda(), torch.tensor(i2).cuda())
    assert np.allclose(K.diagonal().detach().cpu().numpy(), np.sum(phi[i1] * phi[i2], axis=1), rtol=1e-5, atol=1e-5)


def test_diag_on_a_graph_whose_dense_rows_would_not_fit():
    """N = 2^20 ring, 4096 prediction nodes: the densified row sets of the reference's diag branch would be
    2 x 16 GiB; the per-pair kernel needs n x L floats."""
    import torch
    from grf_b200 import engine
    from grf_b200.operators import GRFFeatureOperator

    n = 1 << 20
    idx = np.arange(n)
    adj = sp.csr_matrix((np.ones(2 * n), (np.r_[idx, idx], np.r_[(idx + 1) % n, (idx - 1) % n])), shape=(n, n))
    g = engine.DeviceGraph.laplacian_of(adj, torch.device("cuda", 0))
    phi = engine.build_phi_blocks(g, engine.WalkConfig(20, 0.1, 4, seed=1), transpose=False)
    f = torch.tensor([1.0, 0.5, -0.25, 0.125], device="cuda")
    rows = torch.arange(0, n, n // 4096, device="cuda")[:4096]
    op = GRFFeatureOperator(phi, f)[rows]
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    d = op.row_dots_with(op)
    assert torch.cuda.max_memory_allocated() - base < (64 << 20)
    # against the product route on a few rows: ||Phi_f[r]||^2 = (Phi_f Phi_f^T e_r)[r]
    few = rows[:8]
    e = torch.zeros(n, 8, device="cuda")
    e[few, torch.arange(8, device="cuda")] = 1.0
    k = phi.apply(f, phi.apply_t(f, e), rows=few)
    assert torch.allclose(d[:8], k.diagonal(), rtol=1e-5, atol=1e-6)


def test_bilinear_derivative_is_the_per_length_reduction(setup):
    """Upstream protocol method GRFKernelOperator._bilinear_derivative (what gpytorch's inv_quad_logdet backward
    calls for sparse_grf_kernel.py:51-62): d/df sum(left * (K right)) against float64, and against autograd
    through the same operator."""
    torch = setup["torch"]
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseGRFKernel
    from oracle import grf_oracle as orc

    torch.manual_seed(9)
    kern = SparseGRFKernel(4, setup["ops"]).cuda()
    rng = np.random.default_rng(2)
    i1, i2 = rng.permutation(400)[:130], rng.permutation(400)[:70]
    K = kern(torch.tensor(i1).cuda(), torch.tensor(i2).cuda())
    left = torch.tensor(rng.standard_normal((130, 6)).astype(np.float32)).cuda()
    right = torch.tensor(rng.standard_normal((70, 6)).astype(np.float32)).cuda()
    (got,) = K._bilinear_derivative(left, right)
    f = kern.modulator_vector.detach().cpu().numpy()
    want = orc.phi_fgrad_f64([m.astype(np.float32) for m in setup["pp"].step_matrices_scipy], f,
                             left.cpu().numpy(), right.cpu().numpy(), x1=i1, x2=i2)
    assert np.allclose(got.cpu().numpy(), want, rtol=2e-4, atol=2e-4 * np.abs(want).max())
    (left * (K @ right)).sum().backward()
    auto = kern.raw_modulator_vector.grad.cpu().numpy()
    assert np.allclose(got.cpu().numpy(), auto, rtol=1e-4, atol=1e-4 * np.abs(auto).max())
    (got1,) = K._bilinear_derivative(left[:, 0], right[:, 0])          # vectors, as upstream may pass them
    want1 = orc.phi_fgrad_f64([m.astype(np.float32) for m in setup["pp"].step_matrices_scipy], f,
                              left[:, :1].cpu().numpy(), right[:, :1].cpu().numpy(), x1=i1, x2=i2)
    assert np.allclose(got1.cpu().numpy(), want1, rtol=2e-4, atol=2e-4 * np.abs(want1).max())


def test_out_of_range_row_ids_raise_like_the_reference(setup):
    """phi[idx] raises IndexError in the reference (sparse_grf_kernel.py:32-41); the kernels would return a zero
    row for an id outside the local rows (right for a shard, wrong on the whole Phi)."""
    torch = setup["torch"]
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseGRFKernel

    kern = SparseGRFKernel(4, setup["ops"]).cuda()
    ok = torch.tensor([0, 399]).cuda()
    kern(ok, ok)
    for bad in ([0, 400], [-1, 3], [5, 2 ** 31 + 7]):
        with pytest.raises(IndexError):
            kern(torch.tensor(bad).cuda(), ok)
        with pytest.raises(IndexError):
            setup["ops"].phi_blocks.plan(kern.modulator_vector.detach(), 4, x1=torch.tensor(bad).cuda())


def test_dict_keyed_experiment_caches_load_as_operators(setup, tmp_path):
    """The experiments pickle {'step_matrices_torch': [scipy CSR, ...], ...}
    (run_scaling_experiment.py:381-397, data_utils.py:334-343) and rebuild operators from it
    (run_scaling_experiment.py:550-562): same thing here, and the kernel built on them equals the one built on
    the preprocessor's own operators."""
    import pickle

    torch = setup["torch"]
    from efficient_graph_gp_sparse.gptorch_kernels_sparse import SparseGRFKernel
    from efficient_graph_gp_sparse.preprocessor import GraphPreprocessor

    mats = setup["pp"].step_matrices_scipy
    path = str(tmp_path / "step_matrices_sparse_n400_seed7.pkl")
    with open(path, "wb") as fh:
        pickle.dump({"step_matrices_torch": mats, "n_nodes": 400, "seed": 7, "method": "sparse", "config": {}}, fh)
    ops = GraphPreprocessor.load_step_operators(path, torch.device("cuda", 0))
    assert len(ops) == 4 and ops[0].sparse_csr_tensor.is_cuda
    again = GraphPreprocessor(setup["adj"], 30, 0.1, 4, 7, load_from_disk=True, cache_filename=path)
    assert all(csr_bits_equal(a, b) for a, b in zip(again.step_matrices_scipy, mats))
    k_new, k_old = SparseGRFKernel(4, ops).cuda(), SparseGRFKernel(4, setup["ops"]).cuda()
    with torch.no_grad():
        k_new.raw_modulator_vector.copy_(k_old.raw_modulator_vector)
    x = torch.arange(0, 400, 7).cuda()
    v = torch.randn(x.numel(), 3, device="cuda")
    assert torch.allclose(k_new(x, x) @ v, k_old(x, x) @ v, rtol=1e-5, atol=1e-6)
