"""BASELINE config 4 on ONE GPU, graph drawn on the device: R-MAT 2^scale nodes, >= `edges` undirected
edges, W walks/node, L lengths.  Prints per-phase device times and, with --stats, the shape of Phi that the
matvec design depends on (column skew, row lengths, union size).

  python profiles/run_cfg4.py [--scale 22] [--edges 70e6] [--L 5] [--W 100] [--t 16] [--stats] [--reps 3]
"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
from grf_b200 import engine, synth, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--edges", type=float, default=70e6)
ap.add_argument("--L", type=int, default=5)
ap.add_argument("--W", type=int, default=100)
ap.add_argument("--t", type=int, default=16)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--stats", action="store_true")
ap.add_argument("--merged", action="store_true", help="also time the merged (union) layout")
ap.add_argument("--matvec-only", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def timed(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


t0 = time.perf_counter()
g, st = synth.rmat_walk_graph(args.scale, int(args.edges), seed=0, device=dev)
torch.cuda.synchronize()
print(f"graph on device in {time.perf_counter() - t0:.2f} s: {st}", flush=True)
cfg = engine.WalkConfig(args.W, 0.1, args.L, seed=42)
for rep in range(args.reps):
    phi, ms = timed(lambda: engine.build_phi_blocks(g, cfg, transpose=False))
    visits = int(phi.visits)
    print(f"rep{rep}: walk + compaction {ms:.2f} ms; {visits / 1e9:.3f} G walk-steps -> {visits / ms / 1e6:.2f} G/s; "
          f"nnz(Phi)={phi.nnz / 1e6:.1f} M; peak mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB", flush=True)
    del phi
for rep in range(args.reps):
    phi, ms = timed(lambda: engine.build_phi_blocks(g, cfg, transpose=True))
    print(f"rep{rep}: Phi build incl. Phi^T ({len(phi.tblocks)} row blocks): {ms:.2f} ms; "
          f"peak mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB", flush=True)
    if rep + 1 < args.reps:
        del phi
n, L, t = phi.n_rows, args.L, args.t
f = torch.randn(L, device=dev)
v = torch.randn(n, t, device=dev)
out = torch.empty_like(v)
plan = phi.plan(f, t, merged=False)
for _ in range(2):
    plan(v, out)
_, ms = timed(lambda: plan(v, out), 5)
nb = 2 * phi.nnz * 8 + 2 * L * (n + 1) * 4 + 4 * n * t * 4
print(f"matvec t={t} per-length blocks: {ms:.3f} ms -> {nb / ms / 1e6:.0f} GB/s algorithmic "
      f"({nb / ms / 1e6 / 6554.6:.3f} of 6554.6)", flush=True)
print(f"   gather variants chosen (Phi^T V, Phi U): stream={phi.stream_gather}", flush=True)
for mode in ((False, False), (True, True)):
    plan._gt, plan._gf = (32 if mode[0] else 0), (32 if mode[1] else 0)
    _, ms1 = timed(lambda: plan._call(v, None, 1), 5)
    _, ms2 = timed(lambda: plan._call(None, out, 2), 5)
    print(f"   halves with stream={mode[0]}: Phi^T V {ms1:.3f} ms, Phi U {ms2:.3f} ms", flush=True)
if args.merged and len(phi.tblocks) == 1:
    _, ms = timed(lambda: phi.build_union())
    print(f"union build {ms:.1f} ms; nnz_union={phi.nnz_union / 1e6:.1f} M of {phi.nnz / 1e6:.1f} M", flush=True)
    mplan = phi.plan(f, t, merged=True)
    for _ in range(2):
        mplan(v, out)
    _, ms = timed(lambda: mplan(v, out), 5)
    print(f"matvec t={t} merged Phi_f: {ms:.3f} ms -> {nb / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    _, ms1 = timed(lambda: mplan._call(v, None, 1), 5)
    _, ms2 = timed(lambda: mplan._call(None, out, 2), 5)
    print(f"   merged halves: Phi^T V {ms1:.3f} ms, Phi U {ms2:.3f} ms", flush=True)
    if args.stats:
        for name, e in (("Phi_f", mplan.phi.entries), ("Phi_f^T", mplan.phi.tentries)):
            c = e[:, 0] & ((1 << 27) - 1)
            pair = ((c[1:] == c[:-1] + 1) & ((c[:-1] & 1) == 0)).sum().item()
            near = ((c[1:] >> 1) == (c[:-1] >> 1)).sum().item()
            print(f"   merged {name}: {pair / c.numel():.3f} of the entries are followed by their line partner "
                  f"(col ^ 1); same 128-B line as the predecessor: {near / c.numel():.3f}", flush=True)
if args.stats:
    ent = phi.entries
    cols = (ent[:, 0] & ((1 << 27) - 1)).to(torch.int64)
    cnt = torch.bincount(cols, minlength=phi.n_cols)
    srt = cnt.sort(descending=True).values.cumsum(0).double() / max(1, phi.nnz)
    for k in (256, 2048, 16384, 131072, 1 << 20):
        if k <= srt.numel():
            print(f"   top-{k} columns hold {float(srt[k - 1]):.3f} of the entries")
    print(f"   non-empty columns: {int((cnt > 0).sum())}; longest column {int(cnt.max())}")
    rl = (phi.blk_ptr[L::L] - phi.blk_ptr[:-1:L]).double()
    q = torch.tensor([0.5, 0.9, 0.99, 1.0], device=dev, dtype=torch.float64)
    print(f"   Phi row length: mean {float(rl.mean()):.1f}, quantiles 50/90/99/100 % = {torch.quantile(rl[:1 << 24], q).tolist()}")
    cl = cnt.double()
    print(f"   Phi^T blocks: {[(tb.r0, tb.n_rows, tb.nnz) for tb in phi.tblocks]}")
    print(f"   Phi^T row length: mean {float(cl.mean()):.1f}, quantiles = {torch.quantile(cl[:1 << 24], q).tolist()}")
    step = (ent[:, 0] >> 27) & 31
    print(f"   entries per length: {torch.bincount(step.to(torch.int64), minlength=L).tolist()}")
