"""BASELINE config 5 shape: Thompson-sampling BO on a social-shaped R-MAT graph (500 k nodes, 5 M edges),
W walks/node, L = 3 (graph_bo/configs/default_config.yaml:14-31): 100 initial points, batch 50, 10 iterations,
Thompson n_samples = 1 (bo_utils.py:257-276), objective = standardised node degree (database.py:214).

  python profiles/run_bo.py [log2_nodes=19] [edges_millions=5] [W=1000] [iterations=10]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200"), os.path.join(ROOT, "tests")]
import numpy as np, scipy.sparse as sp, torch
from efficient_graph_gp_sparse.preprocessor import GraphPreprocessor
from efficient_graph_gp_sparse.models import SparseGraphGP
from grf_b200.gp_compat import GaussianLikelihood, settings
from test_gpu_walker import _rmat

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 19
m = int(float(sys.argv[2]) * 1e6) if len(sys.argv) > 2 else 5_000_000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
settings.cg_tolerance._global_value = 1e-2

adj = _rmat(scale, m)
n = adj.shape[0]
deg = np.diff(adj.indptr).astype(np.float64)
objective = (deg - deg.mean()) / deg.std()
print(f"graph N={n}, edges={adj.nnz//2}, max degree {int(deg.max())}", flush=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
pp = GraphPreprocessor(adj, walks_per_node=W, p_halt=0.1, max_walk_length=3, random_walk_seed=42, use_tqdm=False)
ops = pp.preprocess_graph()
torch.cuda.synchronize(); t1 = time.perf_counter()
blocks = ops.phi_blocks
print(f"preprocess_graph: {t1-t0:.2f} s ({int(pp._steps_device.visits)/1e9:.2f} G walk-steps, nnz(Phi)={blocks.nnz/1e6:.0f} M)", flush=True)

rng = np.random.default_rng(0)
observed = rng.permutation(n)[:100]
lik = GaussianLikelihood()
lik.noise = 0.01
torch.manual_seed(0)
best = []
times = []
for it in range(iters):
    torch.cuda.synchronize(); ta = time.perf_counter()
    x = torch.tensor(observed, dtype=torch.float32)[:, None].cuda()
    y = torch.tensor(objective[observed], dtype=torch.float32).cuda()
    model = SparseGraphGP(x, y, lik, ops, 3).cuda()
    with torch.no_grad():
        model.covar_module.raw_modulator_vector.copy_(torch.tensor([1.0, -0.5, 0.125]))
    mask = np.ones(n, dtype=bool); mask[observed] = False
    cand = np.flatnonzero(mask)
    sample, info = model.predict(torch.tensor(cand).cuda(), n_samples=1, return_info=True)
    top = torch.topk(sample[0], 50).indices.cpu().numpy()
    observed = np.r_[observed, cand[top]]
    torch.cuda.synchronize(); times.append(time.perf_counter() - ta)
    best.append(float(objective[observed].max()))
    print(f"  BO iteration {it}: {1e3*times[-1]:.1f} ms (CG iterations {info['iterations']}, n_train {x.shape[0]}), best so far {best[-1]:.2f} "
          f"(global max {objective.max():.2f})", flush=True)
print(f"BO loop: median {1e3*np.median(times):.1f} ms per iteration")
