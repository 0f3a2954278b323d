"""The peer-memory exchange alone: U of N x 16 floats summed over the ranks of one box (torchrun).
  python -m torch.distributed.run --nproc-per-node 8 profiles/prof_exchange.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
import torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from grf_b200 import engine, sharding
n = 1 << 22
ex = sharding.make_exchange(n, 16, dev, True)
ex.u.fill_(float(rank + 1))
st = engine._stream(dev)
for _ in range(3):
    ex.reduce(st)
torch.cuda.synchronize(); dist.barrier()
evs = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ex.reduce(st); b.record(); evs.append((a, b))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in evs)
t = torch.tensor([ms[len(ms) // 2]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    w = dist.get_world_size()
    gb = n * 16 * 4 * (w - 1) / w / 1e9
    print(f"{ex.describe()}: {float(t):.3f} ms median (max over ranks) for {n * 64 / 1e6:.0f} MB of U -> {gb / float(t) * 1e3:.0f} GB/s per direction and rank")
dist.destroy_process_group()
