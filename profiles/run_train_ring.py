"""BASELINE config 3: sparse GRF GP regression on a ring of N = 2^log2_n nodes -- Adam on -mll/n
(modulator + noise), CG with `probes` probe vectors, as run_scaling_experiment.py:53-116,596-623
(lr 0.1, 50 epochs, W=100, p_halt=0.1, L=3, 60 % train, cg_tolerance 1e-2).

  python profiles/run_train_ring.py [log2_n=20] [probes=16] [epochs=50]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import numpy as np, scipy.sparse as sp, torch
from efficient_graph_gp_sparse.preprocessor import GraphPreprocessor
from efficient_graph_gp_sparse.models import SparseGraphGP
from grf_b200.gp_compat import GaussianLikelihood
from grf_b200.mll import neg_mll_backward

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
probes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 50
n = 1 << scale
i = np.arange(n)
adj = sp.csr_matrix((np.ones(2 * n), (np.r_[i, i], np.r_[(i + 1) % n, (i - 1) % n])), shape=(n, n))
np.random.seed(42)
th = 2 * np.pi * i / n
y = 2 * np.sin(2 * th) + 0.5 * np.cos(4 * th) + 0.3 * np.sin(th) + 0.1 * np.random.randn(n)
perm = np.random.permutation(n)
train, test = np.sort(perm[: int(0.6 * n)]), np.sort(perm[int(0.6 * n):int(0.8 * n)])

torch.cuda.synchronize(); t0 = time.perf_counter()
pp = GraphPreprocessor(adj, walks_per_node=100, p_halt=0.1, max_walk_length=3, random_walk_seed=42, use_tqdm=False)
ops = pp.preprocess_graph()
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"N={n}: preprocess_graph (Laplacian + walks + step matrices to host + Phi blocks) {t1-t0:.2f} s")
torch.manual_seed(42)
lik = GaussianLikelihood()
xt = torch.tensor(train, dtype=torch.float32)[:, None].cuda()
yt = torch.tensor(y[train], dtype=torch.float32).cuda()
model = SparseGraphGP(xt, yt, lik, ops, 3).cuda()
lik = lik.cuda()
opt = torch.optim.Adam(list(model.parameters()) + list(lik.parameters()), lr=0.1)
gen = torch.Generator(device="cuda").manual_seed(0)
times = []
for ep in range(epochs):
    torch.cuda.synchronize(); ta = time.perf_counter()
    opt.zero_grad()
    info = neg_mll_backward(model.covar_module, lik, xt, yt, num_probes=probes, cg_tolerance=1e-2, generator=gen)
    opt.step()
    torch.cuda.synchronize(); times.append(time.perf_counter() - ta)
    if ep % 10 == 0 or ep == epochs - 1:
        print(f"  epoch {ep:3d}: loss {info['loss']:.4f}, {1e3*times[-1]:.1f} ms, CG iterations {info['cg_iterations']}, noise {float(lik.noise):.4f}, "
              f"modulator {model.covar_module.modulator_vector.detach().cpu().numpy().round(3)}")
print(f"training: {epochs} epochs, median {1e3*np.median(times):.1f} ms/epoch, total {sum(times):.2f} s")
with torch.no_grad():
    mean = model.posterior_mean(torch.tensor(test).cuda(), cg_tolerance=1e-3)
rmse = float(torch.sqrt(torch.mean((mean.cpu() - torch.tensor(y[test], dtype=torch.float32)) ** 2)))
print(f"test RMSE {rmse:.4f} (noise std of the data 0.1)")
