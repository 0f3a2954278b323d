"""Where does the end-to-end (host CSR in -> host scipy CSR list out) time go?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import numpy as np, torch
import bench
from grf_b200 import engine, _lib

lap = bench.grid_laplacian(316, 316)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T(); g = engine.DeviceGraph.from_scipy(lap)
    t1 = T(); cfg = engine.WalkConfig(100, 0.1, 5, seed=42); steps = engine.build_step_matrices(g, cfg)
    t2 = T(); off = steps.offsets.cpu().numpy(); col = steps.col.cpu().numpy(); val = steps.val.cpu().numpy()
    t3 = T(); mats = steps.to_scipy()
    t4 = T()
    print(f"it{it}: H2D graph {1e3*(t1-t0):.2f} ms | walk+compact {1e3*(t2-t1):.2f} | raw D2H {1e3*(t3-t2):.2f} | to_scipy (incl. D2H again) {1e3*(t4-t3):.2f}")
