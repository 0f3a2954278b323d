"""Run the walker (+ compaction) for BASELINE config 2 a few times (used under ncu and for event timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
import bench
from grf_b200 import engine

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
W = int(sys.argv[2]) if len(sys.argv) > 2 else 100
L = int(sys.argv[3]) if len(sys.argv) > 3 else 5
lap = bench.grid_laplacian(316, 316)
g = engine.DeviceGraph.from_scipy(lap)
cfg = engine.WalkConfig(W, 0.1, L, seed=42)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for i in range(reps + 2):
    visits = torch.zeros(1, dtype=torch.int64, device="cuda")
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); st = engine.run_walker(g, cfg, visits=visits); b.record()
    torch.cuda.synchronize()
    if i >= 2:
        times.append(a.elapsed_time(b) * 1e3)
    v = int(visits)
    del st
times.sort()
print(f"W={W} L={L} visits={v} walker us: min {times[0]:.1f} median {times[len(times)//2]:.1f} -> {v/times[len(times)//2]*1e-3:.2f} G walk-steps/s")
