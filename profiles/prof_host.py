"""Host-side cost of the Phi build phases at BASELINE config 2: wall time the Python thread spends
issuing each phase (no synchronisation inside) next to the device time of the same phase.

  python profiles/prof_host.py [world=1]

world > 1 emulates rank 0 of a row-sharded run on ONE GPU (graph of world x the nodes, the first
1/world of the start nodes, no process group): what the global column count costs a shard.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
import bench
from grf_b200 import engine, _lib

dev = torch.device("cuda:0")
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lap = bench.grid_laplacian(bench.GRID_NX, bench.GRID_NY * world)
g = engine.DeviceGraph.from_scipy(lap, dev)
cfg = engine.WalkConfig(bench.W, bench.P_HALT, bench.L, seed=bench.SEED)
f = torch.randn(bench.L, device=dev)
n = g.n_nodes // world
acc = {}


def phase(name, fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    out = fn()
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    h, d = acc.setdefault(name, ([], []))
    h.append((t1 - t0) * 1e6)
    d.append(a.elapsed_time(b) * 1e3)
    return out


for it in range(12):
    st = phase("walk", lambda: engine.run_walker(g, cfg, 0, n, count_columns=True))
    phi = phase("compact", lambda: engine._blocks_from_staging(st, cfg, g.n_nodes, _lib.SCALE_MUL_RECIP))
    del st
    phase("transpose", lambda: phi.build_transpose())
    plan = phase("plan", lambda: phi.plan(f, 16, merged=False))
    v = torch.randn(n, 16, device=dev)
    out = torch.empty_like(v)
    phase("matvec", lambda: plan(v, out))
    del phi, plan
for k, (h, d) in acc.items():
    h, d = sorted(h[2:]), sorted(d[2:])
    print(f"{k:10s} host issue {h[len(h)//2]:7.1f} us   device {d[len(d)//2]:7.1f} us")
