import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch, numpy as np
from grf_b200 import engine, synth
dev = torch.device("cuda", 0)
adj = synth.rmat_adjacency(22, 70_000_000, seed=0, device=dev).to_scipy()
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); g = engine.DeviceGraph(adj.indptr, adj.indices, adj.data, adj.shape[0], dev); torch.cuda.synchronize(); t1 = time.perf_counter()
    lap = g.laplacian(); torch.cuda.synchronize(); t2 = time.perf_counter()
    c = adj.has_canonical_format; t3 = time.perf_counter()
    print(f"threads={engine._UPLOAD_THREADS} upload {1e3*(t1-t0):.1f} ms ({(adj.indptr.nbytes+adj.indices.nbytes+adj.data.nbytes)/(t1-t0)/1e9:.1f} GB/s), laplacian {1e3*(t2-t1):.1f} ms, canonical check {1e3*(t3-t2):.2f} ms", flush=True)
    del g, lap
