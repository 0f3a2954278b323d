"""The oracle is pinned bit-for-bit to outputs of the reference itself
(fixtures made by tests/golden/make_golden.py from /root/reference)."""

import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, golden_csr, load_golden
from oracle import grf_oracle as orc


def _flat_draws(trace, walk_ids, L):
    out = []
    tu, tk = trace
    for g in walk_ids:
        for s in range(L):
            u = tu[g * L + s]
            if not np.isnan(u):
                out.append((0, float(u)))
                k = tk[g * L + s]
                if k >= 0:
                    out.append((1, float(k)))
    return np.array(out, dtype=np.float64).reshape(-1, 2)


def _check_draws(z, key, got):
    """Small fixtures store the reference's PCG64 draw stream whole, large ones its head, length and sha256."""
    if key in z.files:
        assert np.array_equal(got, z[key])
        return
    import hashlib

    assert got.shape[0] == int(z[key + "_count"])
    assert np.array_equal(got[:64], z[key + "_head"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, z[key + "_sha256"])


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        got = orc.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert [int(x) for x in got] == want


SPARSE = sorted(glob.glob(os.path.join(GOLDEN, "sparse_*.npz")))
DENSE = sorted(glob.glob(os.path.join(GOLDEN, "dense_*.npz")))


@pytest.mark.parametrize("path", SPARSE, ids=[os.path.basename(p)[:-4] for p in SPARSE])
def test_sparse_sampler_matches_reference(path):
    z = np.load(path)
    W, p, L = int(z["W"]), float(z["p_halt"]), int(z["L"])
    seed = None if int(z["seed"]) < 0 else int(z["seed"])
    nproc = int(z["n_processes"])
    graph = golden_csr(z, "graph")
    mats, trace = orc.sparse_step_matrices(graph, W, p, L, seed=seed, n_processes=nproc, record=True)
    n = graph.shape[0]
    for s in range(L):
        want = golden_csr(z, f"step{s}", shape=(n, n))
        assert np.array_equal(mats[s].indptr, want.indptr)
        assert np.array_equal(mats[s].indices, want.indices)
        assert np.array_equal(mats[s].data, want.data)      # bit-exact float64
        assert int(z[f"step{s}_sorted"]) == 1
    # the draws the oracle consumed == the PCG64 draws the reference consumed
    for i, chunk in enumerate(np.array_split(np.arange(n), nproc)):
        ids = [int(c) * W + w for c in chunk for w in range(W)]
        _check_draws(z, f"draws{i}", _flat_draws(trace, ids, L))
    # replaying the recorded trace reproduces the same matrices
    again = orc.step_matrices_from_draws(graph, W, p, L, orc.TraceDraws(*trace))
    for s in range(L):
        assert np.array_equal(again[s].data, mats[s].data) and np.array_equal(again[s].indices, mats[s].indices)


@pytest.mark.parametrize("path", DENSE, ids=[os.path.basename(p)[:-4] for p in DENSE])
def test_dense_sampler_matches_reference(path):
    z = np.load(path)
    W, p, L = int(z["W"]), float(z["p_halt"]), int(z["L"])
    seed = None if int(z["seed"]) < 0 else int(z["seed"])
    nproc = int(z["n_processes"])
    got, trace = orc.dense_step_tensor(z["graph"], W, p, L, seed=seed, n_processes=nproc,
                                       ablation=bool(int(z["ablation"])), record=True)
    assert np.array_equal(got, z["tensor"])
    n = z["graph"].shape[0]
    if nproc == 1 or n < 2 * nproc:
        chunks = [np.arange(n)]
    else:
        chunks = np.array_split(np.arange(n), nproc)
    for i, chunk in enumerate(chunks):
        ids = [int(c) * W + w for c in chunk for w in range(W)]
        assert np.array_equal(_flat_draws(trace, ids, L), z[f"draws{i}"])


def test_laplacians_and_kernels_match_reference():
    z = load_golden("kernels.npz")
    nproc = int(z["n_processes"])
    for name in ("cycle4", "grid6x4", "gnm40w"):
        adj = z[name + "_adj"]
        lap_s = orc.normalized_laplacian_sparse(sp.csr_matrix(adj))
        want = golden_csr(z, name + "_lap_sparse")
        assert np.array_equal(lap_s.indptr, want.indptr) and np.array_equal(lap_s.indices, want.indices)
        assert np.array_equal(lap_s.data, want.data)
        assert np.array_equal(orc.normalized_laplacian_dense(adj), z[name + "_lap_dense"])
        f = z[name + "_f"]
        ks = orc.grf_kernel_sparse(sp.csr_matrix(adj), f, 10, 0.2, 3, n_processes=nproc).toarray()
        assert np.array_equal(ks, z[name + "_K_sparse"])
        kd = orc.grf_kernel_dense(adj, f, 10, 0.2, 3, n_processes=nproc)
        assert np.allclose(kd, z[name + "_K_dense"], rtol=0, atol=1e-12)  # BLAS dgemm order may differ


def test_reference_test_invariants():
    """What the reference's own tests pin (tests/test_grf_*.py): shapes,
    M_0 = I, K symmetric PSD on the 4-cycle."""
    cyc = load_golden("kernels.npz")["cycle4_adj"]
    mats = orc.sparse_step_matrices(sp.csr_matrix(cyc), 5, 0.2, 3, seed=0, n_processes=1)
    assert len(mats) == 3 and all(m.shape == (4, 4) for m in mats)
    assert np.allclose(mats[0].diagonal(), 1.0, atol=1e-6)
    k = orc.grf_kernel_sparse(sp.csr_matrix(cyc), [1.0, 0.5, 0.25], 10, 0.2, 3).toarray()
    assert np.allclose(k, k.T, atol=1e-8) and np.linalg.eigvalsh(k).min() >= -1e-8


def test_estimator_is_unbiased_for_matrix_powers():
    """E[M_l] = A^l for the walk graph A; loose statistical check."""
    cyc = load_golden("kernels.npz")["cycle4_adj"]
    lap = orc.normalized_laplacian_sparse(sp.csr_matrix(cyc))
    mats = orc.step_matrices_from_draws(lap, 1500, 0.1, 3, orc.PhiloxDraws(1234))
    a = lap.toarray()
    assert np.allclose(mats[1].toarray(), a, atol=0.12)
    assert np.allclose(mats[2].toarray(), a @ a, atol=0.2)


@pytest.mark.parametrize("path", SPARSE, ids=[os.path.basename(p)[:-4] for p in SPARSE])
def test_cpu_baseline_port_matches_reference(path):
    """The timed CPU baseline (oracle/cpu_baseline.py, fork pool) is the reference's algorithm bit for bit."""
    from oracle import cpu_baseline

    z = np.load(path)
    W, p, L = int(z["W"]), float(z["p_halt"]), int(z["L"])
    seed = None if int(z["seed"]) < 0 else int(z["seed"])
    graph = golden_csr(z, "graph")
    mats, visits = cpu_baseline.sampler_pool(graph, W, p, L, seed=seed, n_processes=int(z["n_processes"]),
                                             return_visits=True)
    n = graph.shape[0]
    for s in range(L):
        want = golden_csr(z, f"step{s}", shape=(n, n))
        assert np.array_equal(mats[s].indptr, want.indptr) and np.array_equal(mats[s].indices, want.indices)
        assert np.array_equal(mats[s].data, want.data)
    assert visits >= n * W


SLICES = sorted(glob.glob(os.path.join(GOLDEN, "slice_*.npz")))


@pytest.mark.parametrize("path", SLICES, ids=[os.path.basename(p)[:-4] for p in SLICES])
def test_worker_slice_matches_reference(path):
    """One reference worker on a row slice of a 2^20 .. 2^25-node ring (the fixtures behind the GPU walker's
    64-bit-key and slice-local-trace replay tests): rows and draw stream, bit for bit."""
    z = np.load(path)
    n, lo, hi = 1 << int(z["log2_n"]), int(z["lo"]), int(z["hi"])
    W, p, L = int(z["W"]), float(z["p_halt"]), int(z["L"])
    lap = orc.ring_laplacian_csr(n, float(z["lap_diag"]), float(z["lap_off"]))
    mats, trace = orc.worker_slice_rows(lap.indptr, lap.indices, lap.data, n, lo, hi, W, p, L,
                                        int(z["worker_seed"]), record=True)
    for s in range(L):
        assert np.array_equal(mats[s].indptr, z[f"step{s}_indptr"])
        assert np.array_equal(mats[s].indices, z[f"step{s}_indices"])
        assert np.array_equal(mats[s].data, z[f"step{s}_data"])
    _check_draws(z, "draws", _flat_draws(trace, range((hi - lo) * W), L))


def test_matvec_layer_matches_reference():
    """M1 / P1 pinned: the oracle's restatement of graph_preprocessor.py:117-139 (float32 / int64 conversion),
    sparse_lo.py:16-25 (M_l X, M_l^T X through .t().to_sparse_csr()) and the per-length scale + sum of
    sparse_grf_kernel.py:51-62, against outputs of the reference's own classes (torch-CPU sparse CSR)."""
    import torch

    z = load_golden("matvec_layer.npz")
    W, p, L, nproc = int(z["W"]), float(z["p_halt"]), int(z["L"]), int(z["n_processes"])
    adj = golden_csr(z, "adj")
    n = adj.shape[0]
    lap = orc.normalized_laplacian_sparse(adj)
    mats = orc.sparse_step_matrices(lap, W, p, L, seed=int(z["seed"]), n_processes=nproc)
    X = z["X"]
    for s in range(L):
        want = golden_csr(z, f"step{s}", shape=(n, n))
        assert np.array_equal(mats[s].indptr, want.indptr) and np.array_equal(mats[s].indices, want.indices)
        assert np.array_equal(mats[s].data, want.data)
        t = orc.torch_csr_f32(mats[s])
        assert np.array_equal(t.crow_indices().numpy(), z[f"torch{s}_crow"])
        assert np.array_equal(t.col_indices().numpy(), z[f"torch{s}_col"])
        assert np.array_equal(t.values().numpy().view(np.int32), z[f"torch{s}_val"].view(np.int32))
        assert np.array_equal(t.matmul(torch.from_numpy(X)).numpy(), z[f"matmul{s}"])
        tt = t.t().to_sparse_csr()
        assert np.array_equal(tt.col_indices().numpy(), z[f"torchT{s}_col"])
        assert np.array_equal(tt.matmul(torch.from_numpy(X)).numpy(), z[f"tmatmul{s}"])
    got = orc.phi_matvec_reference_torch(mats, z["f"], z["V"])
    assert np.allclose(got, z["K_V"], rtol=0, atol=2e-6 * np.abs(z["K_V"]).max())   # same ops; sum order of the L terms
    got = orc.phi_matvec_reference_torch(mats, z["f"], z["V2"], x1=z["x1"], x2=z["x2"])
    assert np.allclose(got, z["K_x1x2_V2"], rtol=0, atol=2e-6 * np.abs(z["K_x1x2_V2"]).max())
    # and the float64 evaluation the GPU tests compare against agrees with the reference to fp32 round-off
    f64 = orc.phi_matvec_f64([m.astype(np.float32) for m in mats], z["f"], z["V"])
    assert np.abs(f64 - z["K_V"]).max() <= 2e-5 * np.abs(f64).max()
