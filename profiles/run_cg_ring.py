"""BASELINE config 3 shape: ring graph N=2^20, W=100, L=3, 60 % train nodes, CG on (K + s2 I) [y | 16 probes].

  python profiles/run_cg_ring.py [log2_n=20] [t=17]
Data as the reference's scaling experiment (run_scaling_experiment.py:157-176): y = 2 sin 2th + 0.5 cos 4th +
0.3 sin th + N(0, 0.1^2), np.random.seed(42), random 60 % train split; CG tolerance 1e-2, no preconditioner."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import numpy as np, scipy.sparse as sp, torch
from grf_b200 import engine
from grf_b200.cg import linear_cg, linear_cg_fused
from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
t = int(sys.argv[2]) if len(sys.argv) > 2 else 17
n = 1 << scale
i = np.arange(n)
adj = sp.csr_matrix((np.ones(2 * n), (np.r_[i, i], np.r_[(i + 1) % n, (i - 1) % n])), shape=(n, n))
lap = get_normalized_laplacian(adj)
np.random.seed(42)
th = 2 * np.pi * i / n
y = 2 * np.sin(2 * th) + 0.5 * np.cos(4 * th) + 0.3 * np.sin(th) + 0.1 * np.random.randn(n)
train = np.sort(np.random.permutation(n)[: int(0.6 * n)])

g = engine.DeviceGraph.from_scipy(lap)
torch.cuda.synchronize(); t0 = time.perf_counter()
phi = engine.build_phi_blocks(g, engine.WalkConfig(100, 0.1, 3, seed=42))
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"N={n} Phi build {1e3*(t1-t0):.1f} ms (first call), nnz={phi.nnz}, walk-steps={int(phi.visits)}")
torch.manual_seed(42)
f = torch.randn(3).cuda()
x = torch.tensor(train).cuda()
rhs = torch.cat([torch.tensor(y[train], dtype=torch.float32)[:, None], torch.randn(train.size, t - 1)], 1).cuda()
sigma2 = 0.1
for merged in (True, False):
    plan = phi.plan(f, t, x1=x, x2=x, merged=merged)
    out = torch.empty_like(rhs)
    for _ in range(3): plan(rhs, out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): plan(rhs, out)
    b.record(); torch.cuda.synchronize()
    print(f"  K[x_train, x_train] @ [n_train={train.size}, t={t}] matvec ({'merged' if merged else 'per-length'}): {a.elapsed_time(b)/20*1e3:.1f} us")
plan = phi.plan(f, t, x1=x, x2=x)
for name, fn in (("fused", lambda: linear_cg_fused(plan, rhs, sigma2, tolerance=1e-2, return_info=True)),
                 ("torch-op", lambda: linear_cg(lambda v: plan(v) + sigma2 * v, rhs, tolerance=1e-2, return_info=True))):
    fn()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    sol, info = fn()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    res = (plan(sol) + sigma2 * sol - rhs).norm(dim=0) / rhs.norm(dim=0)
    print(f"  CG {name}: {info['iterations']} iterations, {1e3*(t3-t2):.2f} ms total, {1e6*(t3-t2)/info['iterations']:.1f} us/iteration, "
          f"max relative residual {float(res.max()):.2e}")
