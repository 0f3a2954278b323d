"""Build Phi for BASELINE config 2 and run a few kernel matvecs (used under ncu and for event timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
import bench
from grf_b200 import engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
t = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lap = bench.grid_laplacian(316, 316)
g = engine.DeviceGraph.from_scipy(lap)
phi = engine.build_phi_blocks(g, engine.WalkConfig(100, 0.1, 5, seed=42))
torch.manual_seed(42)
f = torch.randn(5).cuda()
v = torch.randn(phi.n_rows, t, device="cuda")
out = torch.empty_like(v)
phi.use_tiles = bool(int(os.environ.get('GRF_TILES', '0')))
plan = phi.plan(f, t)          # merged Phi_f on the union pattern (default)
print('nnz per-length', phi.nnz, 'nnz union', phi.nnz_union)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    plan(v, out)
torch.cuda.synchronize()
times = []
for _ in range(reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan(v, out); b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b) * 1e3)
times.sort()
print(f"nnz={phi.nnz} t={t} matvec us: min {times[0]:.1f} median {times[len(times)//2]:.1f} max {times[-1]:.1f}")
# warm (no flush) back-to-back, as inside a CG loop
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    plan(v, out)
b.record(); torch.cuda.synchronize()
print(f"back-to-back (L2 warm) us per matvec: {a.elapsed_time(b)*1e3/50:.1f}")
plan2 = phi.plan(f, t, merged=False)
for _ in range(3):
    plan2(v, out)
torch.cuda.synchronize()
times = []
for _ in range(reps):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan2(v, out); b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b) * 1e3)
times.sort()
print(f"  per-length blocks (f applied per entry): matvec us: min {times[0]:.1f} median {times[len(times)//2]:.1f} max {times[-1]:.1f}")
