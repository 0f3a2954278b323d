"""Tiny pass through every kernel of the library (run under compute-sanitizer --tool memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200"), os.path.join(ROOT, "tests")]
import numpy as np, scipy.sparse as sp, torch
from grf_b200 import engine, _lib
from grf_b200.cg import linear_cg_fused
from gpu_util import grid_graph, powerlaw_graph, random_graph, ring_graph
from efficient_graph_gp_sparse.preprocessor import GraphPreprocessor
from efficient_graph_gp_sparse.models import SparseGraphGP
from grf_b200.gp_compat import GaussianLikelihood
from grf_b200.mll import neg_mll_backward

torch.manual_seed(0)
for name, adj, W, L in [("grid", grid_graph(23, 17), 20, 4), ("powerlaw", powerlaw_graph(1500, 12000, 1), 40, 3),
                        ("iso", random_graph(200, 90, 2, weighted=True), 7, 5), ("ring", ring_graph(333), 300, 3),
                        ("empty", sp.csr_matrix((40, 40)), 5, 3)]:
    g = engine.DeviceGraph.laplacian_of(adj)
    cfg = engine.WalkConfig(W, 0.1, L, seed=3)
    steps = engine.build_step_matrices(g, cfg, max_stage_bytes=200_000)
    mats = steps.to_scipy()
    phi = engine.build_phi_blocks(g, cfg)
    n = phi.n_rows
    f = torch.randn(L).cuda()
    for t in (1, 3, 16, 17, 40):
        v = torch.randn(n, t).cuda()
        a = phi.matvec(f, v)
        b = phi.plan(f, t)(v)
        x = torch.randperm(n)[: max(1, n // 3)].cuda()
        c = phi.plan(f, t, x1=x, x2=x)(v[: x.numel()].contiguous())
        phi.fgrad(f, v[: x.numel()], v[: x.numel()], x1=x, x2=x)
    phi.use_tiles = True
    phi.win = None
    phi.build_windows()
    phi.matvec(f, torch.randn(n, 16).cuda())
    x = torch.randperm(n)[: max(2, n // 2)].cuda()
    plan = phi.plan(f, 8, x1=x, x2=x)
    linear_cg_fused(plan, torch.randn(x.numel(), 8).cuda(), 0.5, tolerance=1e-3, max_iter=30)
    torch.cuda.synchronize()
    print("ok", name, steps.nnz_per_step())
adj = grid_graph(15, 11)
pp = GraphPreprocessor(adj, 10, 0.1, 3, use_tqdm=False)
ops = pp.preprocess_graph()
lik = GaussianLikelihood()
xt = torch.arange(0, 165, 2, dtype=torch.float32)[:, None].cuda()
model = SparseGraphGP(xt, torch.randn(xt.numel()).cuda(), lik, ops, 3).cuda()
neg_mll_backward(model.covar_module, lik.cuda(), xt, model.y_train, num_probes=4)
model.predict(torch.arange(1, 165, 2).cuda(), n_samples=3, cg_tolerance=1e-2)
torch.cuda.synchronize()
print("all ok")
