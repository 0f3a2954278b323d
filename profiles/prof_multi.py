"""Where a sharded step spends its time (run under torchrun, 2+ ranks): the one-off preparation of the
matvec plan and the pieces of one sharded product, each timed with CUDA events on every rank.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/prof_multi.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch, torch.distributed as dist
import bench
from grf_b200 import engine, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
lap = bench.grid_laplacian(bench.GRID_NX, bench.GRID_NY * world)
n = lap.shape[0]
per = n // world
lo, hi = rank * per, (rank + 1) * per if rank < world - 1 else n
g = engine.DeviceGraph.from_scipy(lap, dev)
cfg = engine.WalkConfig(bench.W, bench.P_HALT, bench.L, seed=bench.SEED)
f = torch.randn(bench.L, device=dev)
v = torch.randn(hi - lo, 16, device=dev)
out = torch.empty_like(v)
bounds = [r * per for r in range(world)] + [n]
hint = g.shared_columns(bounds, bench.L)
acc = {}


def timed(name, fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); r = fn(); b.record(); t1 = time.perf_counter()
    acc.setdefault(name, []).append((a, b, (t1 - t0) * 1e6))
    return r


for it in range(8):
    st = engine.run_walker(g, cfg, lo, hi, count_columns=True)
    phi = engine._blocks_from_staging(st, cfg, g.n_nodes, 0)
    del st
    phi.row_lo = lo
    phi.shared_hint = hint
    dist.barrier(); torch.cuda.synchronize()
    timed("transpose", lambda: phi.build_transpose())
    timed("long_rows+tcols", lambda: phi.build_long_rows())
    plan = timed("plan (shared columns)", lambda: phi.plan(f, 16, group=True, merged=False))
    dist.barrier(); torch.cuda.synchronize()
    timed("pass A", lambda: plan._call(v, None, 1))
    timed("reduce_shared", lambda: sharding.reduce_shared(plan.u, plan._shared, plan._shared_buf, None))
    timed("pass B", lambda: plan._call(None, out, 2))
    torch.cuda.synchronize()
    del phi, plan
if rank == 0:
    for k, lst in acc.items():
        d = sorted(a.elapsed_time(b) * 1e3 for a, b, _ in lst[2:])
        h = sorted(x for _, _, x in lst[2:])
        print(f"{k:24s} device {d[len(d)//2]:8.1f} us   host issue {h[len(h)//2]:8.1f} us")
    print("shared columns:", None if plan._shared is None else 0) if False else None
dist.destroy_process_group()
