import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import numpy as np, scipy.sparse as sp, torch
from grf_b200 import engine
from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian
scale, m = 20, 16_000_000
n = 1 << scale
rng = np.random.default_rng(0)
src = np.zeros(m, dtype=np.int64); dst = np.zeros(m, dtype=np.int64)
for bit in range(scale):
    r = rng.random(m)
    src |= ((r >= 0.76).astype(np.int64)) << bit
    dst |= (((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).astype(np.int64) << bit
keep = src != dst
lo, hi = np.minimum(src[keep], dst[keep]), np.maximum(src[keep], dst[keep])
key = np.unique(lo * n + hi); lo, hi = key // n, key % n
adj = sp.csr_matrix((np.ones(2 * key.size), (np.r_[lo, hi], np.r_[hi, lo])), shape=(n, n))
g = engine.DeviceGraph.laplacian_of(adj)
phi = engine.build_phi_blocks(g, engine.WalkConfig(100, 0.1, 5, seed=42))
t = (phi.tblk_ptr[1:] - phi.tblk_ptr[:-1]).long()
print("segments", t.numel(), "entries", int(t.sum()))
edges = [0, 1, 8, 32, 256, 1024, 4096, 16384, 65536, 262144, 1 << 22]
for a, b in zip(edges[:-1], edges[1:]):
    sel = (t > a) & (t <= b)
    ln = t[sel].double()
    cost = (ln * torch.log2(ln.clamp(min=2)) ** 2).sum().item() if ln.numel() else 0
    print(f"({a:7d}, {b:7d}]: {int(sel.sum()):9d} segments, {int(ln.sum()):11d} entries, n log^2 n = {cost:.3g}")
print("largest:", torch.topk(t, 8).values.tolist())
