"""C oracle == numpy oracle == reference golden vectors (bit-exact)."""

import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, golden_csr
from oracle import c_oracle, grf_oracle as orc

SPARSE = sorted(glob.glob(os.path.join(GOLDEN, "sparse_*.npz")))


def _same(a, b):
    a, b = a.tocsr(), b.tocsr()
    return (np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
            and np.array_equal(a.data.view(np.int64), b.data.view(np.int64)))


def test_c_philox_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        ctr = rng.integers(0, 2**32, 4, dtype=np.uint64).astype(np.uint32)
        key = rng.integers(0, 2**32, 2, dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(c_oracle.philox(ctr, key), orc.philox4x32_10(ctr, key))


@pytest.mark.parametrize("path", SPARSE, ids=[os.path.basename(p)[:-4] for p in SPARSE])
def test_c_replay_matches_reference(path):
    z = np.load(path)
    W, p, L = int(z["W"]), float(z["p_halt"]), int(z["L"])
    seed = None if int(z["seed"]) < 0 else int(z["seed"])
    graph = golden_csr(z, "graph")
    _, trace = orc.sparse_step_matrices(graph, W, p, L, seed=seed, n_processes=int(z["n_processes"]), record=True)
    mats = c_oracle.step_matrices(graph, W, p, L, draw_mode=c_oracle.DRAW_TRACE, trace=trace)
    n = graph.shape[0]
    for s in range(L):
        assert _same(mats[s], golden_csr(z, f"step{s}", shape=(n, n)))


@pytest.mark.parametrize("load_mode", [0, 1, 2])
def test_c_philox_matches_numpy_walker(load_mode):
    z = np.load(os.path.join(GOLDEN, "sparse_gnm40_weighted_iso_lap_p4.npz"))
    graph = golden_csr(z, "graph")
    want = orc.step_matrices_from_draws(graph, 9, 0.15, 5, orc.PhiloxDraws(77), load_mode=load_mode)
    got, visits = c_oracle.step_matrices(graph, 9, 0.15, 5, seed=77, load_mode=load_mode, return_visits=True)
    for s in range(5):
        assert _same(got[s], want[s])
    assert visits >= 40 * 9
    # row sharding is a pure slicing of the rows
    part = c_oracle.step_matrices(graph, 9, 0.15, 5, seed=77, load_mode=load_mode, start_lo=13, start_hi=29)
    for s in range(5):
        assert _same(part[s], want[s][13:29])
