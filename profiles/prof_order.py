"""Experiment: does handing the rows to the SpMM warps grouped by gather-round count (stable, so
neighbouring rows stay neighbours inside a group) cut the padded slots enough to pay for the lost
locality?  Merged Phi_f at BASELINE config 2, t = 16.

  python profiles/prof_order.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
import bench
from grf_b200 import engine

dev = torch.device("cuda:0")
lap = bench.grid_laplacian(316, 316)
g = engine.DeviceGraph.from_scipy(lap, dev)
cfg = engine.WalkConfig(100, 0.1, 5, seed=42)
phi = engine.build_phi_blocks(g, cfg)
f = torch.randn(5, device=dev)
n = phi.n_rows
v = torch.randn(n, 16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def time_plan(plan, label):
    out = torch.empty(plan.n1, 16, device=dev)
    for _ in range(3):
        plan(v, out)
    ts = []
    for _ in range(15):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan(v, out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{label:40s} min {ts[0]:.1f} median {ts[len(ts)//2]:.1f} us")
    return out


base = phi.plan(f, 16)
ref = time_plan(base, "rows in natural order")
for name, key in (("rounds of 16", lambda ln: (ln + 15) // 16), ("rounds of 8", lambda ln: (ln + 7) // 8), ("exact length", lambda ln: ln)):
    plan = phi.plan(f, 16)
    m = plan.phi
    lens = (m.blk_ptr[1:] - m.blk_ptr[:-1]).long()
    tlens = (m.tblk_ptr[1:] - m.tblk_ptr[:-1]).long()
    perm = torch.sort(key(lens), stable=True).indices.to(torch.int32)
    tperm = torch.sort(key(tlens), stable=True).indices.to(torch.int32)
    plan2 = phi.plan(f, 16, x1=perm)
    plan2.phi._tcols = tperm.contiguous()
    plan2._c = plan2.phi.c_struct(plan2.ldu)
    out = time_plan(plan2, f"grouped by {name}")
    back = torch.empty_like(out)
    back[perm.long()] = out
    print("   max |diff| vs natural order:", float((back - ref).abs().max()))
