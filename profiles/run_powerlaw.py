"""BASELINE config 4 shape on ONE GPU: R-MAT power-law graph, W=100, L=5 (or 3): Phi build time.

  python profiles/run_powerlaw.py [log2_nodes=22] [edges_millions=70] [L=5] [W=100]

R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05), seed 0, symmetrised, de-duplicated, self-loops removed,
unit weights."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import numpy as np, scipy.sparse as sp, torch
from grf_b200 import engine, _lib
from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
m = int(float(sys.argv[2]) * 1e6) if len(sys.argv) > 2 else 70_000_000
L = int(sys.argv[3]) if len(sys.argv) > 3 else 5
W = int(sys.argv[4]) if len(sys.argv) > 4 else 100
n = 1 << scale

def rmat(scale, m, seed=0, a=0.57, b=0.19, c=0.19):
    rng = np.random.default_rng(seed)
    src = np.zeros(m, dtype=np.int64); dst = np.zeros(m, dtype=np.int64)
    for bit in range(scale):
        r = rng.random(m)
        src |= ((r >= a + b).astype(np.int64)) << bit                        # quadrants c, d: src bit set
        dst |= (((r >= a) & (r < a + b)) | (r >= a + b + c)).astype(np.int64) << bit   # quadrants b, d
    return src, dst

t0 = time.perf_counter()
src, dst = rmat(scale, m)
keep = src != dst
lo, hi = np.minimum(src[keep], dst[keep]), np.maximum(src[keep], dst[keep])
key = np.unique(lo * n + hi)
lo, hi = key // n, key % n
adj = sp.csr_matrix((np.ones(2 * key.size), (np.r_[lo, hi], np.r_[hi, lo])), shape=(n, n))
t1 = time.perf_counter()
lap = get_normalized_laplacian(adj)
t2 = time.perf_counter()
deg = np.diff(adj.indptr)
print(f"graph: N={n} undirected edges={key.size} nnz(L)={lap.nnz} max degree={deg.max()} isolated={int((deg==0).sum())} "
      f"| host: generate {t1-t0:.1f}s, Laplacian {t2-t1:.1f}s", flush=True)

torch.cuda.synchronize()
t3 = time.perf_counter()
g = engine.DeviceGraph.from_scipy(lap)
torch.cuda.synchronize()
print(f"H2D graph {time.perf_counter()-t3:.2f}s", flush=True)
cfg = engine.WalkConfig(W, 0.1, L, seed=42)
for rep in range(int(os.environ.get("GRF_REPS", "2"))):
    torch.cuda.synchronize(); t4 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    phi = engine.build_phi_blocks(g, cfg, transpose=False)
    b.record(); torch.cuda.synchronize(); t5 = time.perf_counter()
    visits = int(phi.visits)
    print(f"rep{rep}: Phi build (walk + compaction, {W} walks/node, L={L}): {t5-t4:.3f}s wall, {a.elapsed_time(b)/1e3:.3f}s device; "
          f"{visits/1e9:.3f} G walk-steps -> {visits/(t5-t4)/1e9:.2f} G walk-steps/s; nnz(Phi blocks)={phi.nnz/1e6:.1f} M; "
          f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
    if rep == 0:
        del phi
t6 = time.perf_counter()
phi.build_transpose(); torch.cuda.synchronize()
print(f"Phi^T blocks: {time.perf_counter()-t6:.3f}s", flush=True)
f = torch.randn(L, device="cuda"); v = torch.randn(phi.n_rows, 16, device="cuda"); out = torch.empty_like(v)
plan = phi.plan(f, 16, merged=False)
for _ in range(2): plan(v, out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): plan(v, out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
nb = 2 * phi.nnz * 8 + 2 * L * (phi.n_rows + 1) * 4 + 4 * phi.n_rows * 16 * 4
print(f"matvec t=16 (per-length blocks): {ms:.2f} ms -> {nb/ms/1e6:.0f} GB/s algorithmic", flush=True)
