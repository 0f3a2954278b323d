"""-m gpu tests that need two or more B200s in the box (skipped on a single GPU): the real multi-rank path --
NCCL, symmetric memory, grf_exchange_sum -- against the single-GPU product and the float64 oracle; and the
device guard of the C entry points."""

import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_two_ranks_match_the_single_gpu_product_and_the_oracle():
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "multi_rank_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if res.returncode != 0 or "MULTI_RANK_OK" not in res.stdout:
        lines = [ln for ln in res.stderr.splitlines() if "Error" in ln or "error" in ln][-12:]
        pytest.fail("multi-rank worker failed:\n" + "\n".join(lines) + "\n--- stdout ---\n" + res.stdout[-1500:])


def test_entry_points_run_on_the_device_of_their_stream():
    """building and multiplying on cuda:1 while cuda:0 is current must work (every C entry point
    switches to the device that owns the stream it is given)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch

    from gpu_util import random_graph
    from grf_b200 import engine
    from oracle import grf_oracle

    torch.cuda.set_device(0)
    lap = grf_oracle.normalized_laplacian_sparse(random_graph(400, 1500, 1, weighted=True))
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        g = engine.DeviceGraph.from_scipy(lap, dev)
        phi = engine.build_phi_blocks(g, engine.WalkConfig(20, 0.1, 4, seed=3))
        v = torch.ones(400, 8, device=dev)
        f = torch.tensor([1.0, 0.5, 0.25, 0.125], device=dev)
        outs.append(phi.matvec(f, v).cpu().numpy())
        assert torch.cuda.current_device() == 0
    assert np.array_equal(outs[0], outs[1])
