#!/usr/bin/env python
"""bench.py -- the GRF hot path on B200, measured the way BASELINE.json asks.

Workload (config.workload): BASELINE.json configs[1] -- sparse CSR GRF on a 2-D
grid graph of 316 x 316 = 99 856 nodes, walks_per_node = 100, max_walk_length =
5, p_halt = 0.1, learnable modulator f = randn(5) (torch.manual_seed(42)),
t = 16 right-hand sides.  With N > 1 GPUs the grid grows to 316 x (316 N)
nodes and every rank owns a contiguous block of 99 856 start nodes (weak
scaling; CSR graph replicated; per matvec one all-reduce of the rows of Phi^T V
that more than one rank touches).

One "step" = one pass of the hot path: walker (+ merge) -> compaction into Phi
blocks -> Phi^T blocks (+ the one-off matvec preparation: row census, workspaces)
-> one kernel matvec Phi(Phi^T V).  Every phase is timed
on the device with CUDA events on the launching stream; L2 is flushed (a 256 MB
write) before every timed phase.  `value` = walk-steps executed by all ranks /
max-over-ranks step time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the reference's CPU algorithm (oracle/cpu_baseline.py:
the pure-Python fork-pool sampler, all host cores) on a bounded sample of the
same workload.  oracle/ is used here only as the timed CPU baseline.
"""

from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]

import numpy as np
import scipy.sparse as sp

GRID_NX, GRID_NY = 316, 316
W, P_HALT, L, T_RHS = 100, 0.1, 5, 16
SEED = 42
METRIC = "grf_walk_steps_per_sec"
UNIT = "walk-steps/s"
WALK_BYTES_PER_STEP = 32          # row_ptr pair 8 + col 4 + val 8 (fp64) + merged record 12 (SURVEY 8d, fp64 path)


def grid_laplacian(nx: int, ny: int) -> sp.csr_matrix:
    """Normalized Laplacian of the nx x ny 4-neighbour grid (Kronecker construction as in the reference's
    scalable_bo/bo_utils/data_utils.py:56-60), unit weights."""
    from efficient_graph_gp_sparse.utils_sparse.graph_utils import get_normalized_laplacian

    def path(n):
        return sp.diags([np.ones(n - 1), np.ones(n - 1)], [-1, 1], format="csr")

    adj = (sp.kron(sp.eye(ny), path(nx)) + sp.kron(path(ny), sp.eye(nx))).tocsr()
    return get_normalized_laplacian(adj)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def wait_first_sample(self, timeout=3.0):
        t0 = time.perf_counter()
        while self.proc is not None and time.perf_counter() - t0 < timeout:
            if os.path.getsize(self.out.name) > 0:
                return
            time.sleep(0.01)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.out.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.out.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "warm-up + timed steps + the 20 timed CG matvecs (the timed steps alone last a few ms)"}


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline

    n_gpus = args.gpus
    lap = grid_laplacian(GRID_NX, GRID_NY * n_gpus)
    n = lap.shape[0]
    cores = os.cpu_count() or 1
    n_sample = min(n, 2 * GRID_NX * cores)          # two grid lines of start nodes per core and step
    starts = np.arange(n_sample)
    times, visits = [], 0
    for i in range(args.warmup + args.steps):
        dt, vis = cpu_baseline.time_sampler(lap, W, P_HALT, L, starts, n_processes=cores)
        if i >= args.warmup:
            times.append(dt)
            visits += vis
    total = sum(times)
    value = visits / total
    sample = (f"{n_sample} of {n} start nodes per step (x{W} walks), oracle/cpu_baseline.py fork-pool port of "
              f"sparse_sampler.py:72-132, {cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_gpus, n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus, n_nodes):
    return {
        "workload": "BASELINE.json configs[1]: sparse CSR GRF on 2D grid graph N=100k (316x316 per GPU), "
                    "walks_per_node=100, max_walk_length=5, p_halt=0.1, learnable modulator, t=16",
        "n_nodes": int(n_nodes), "walks_per_node": W, "max_walk_length": L, "p_halt": P_HALT, "rhs_columns": T_RHS,
        "sharding": f"start nodes row-sharded over {n_gpus} GPU(s), CSR graph replicated",
        "l2": "flushed (256 MB write) before every timed phase; staging (480 MB) exceeds L2 as well",
        "host": "Python gc disabled inside the timed region",
    }


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from grf_b200 import _lib, engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lap = grid_laplacian(GRID_NX, GRID_NY * world)
    n = lap.shape[0]
    rows_per = n // world
    lo, hi = rank * rows_per, (rank + 1) * rows_per if rank < world - 1 else n
    graph = engine.DeviceGraph.from_scipy(lap, dev)
    cfg = engine.WalkConfig(W, P_HALT, L, seed=SEED)
    # rows of Phi^T V the ranks must sum: nodes within L - 1 hops of two or more row shards.  A property of
    # the graph and the sharding (like the Laplacian itself), so computed once here, not per step.
    bounds = [r * rows_per for r in range(world)] + [n]
    shared_hint = graph.shared_columns(bounds, L) if world > 1 else None
    torch.manual_seed(42)
    f = torch.randn(L).to(dev)                         # learnable modulator init, sparse_grf_kernel.py:14-17
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    v = torch.randn(hi - lo, T_RHS, device=dev, generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(fn):
        flush.fill_(1)                                  # L2 flush, outside the timed bracket
        a, b = ev(), ev()
        a.record(stream)
        out = fn()
        b.record(stream)
        return out, (a, b)

    lib = _lib.lib()
    stride = lib.grf_walk_stage_stride(W, L)

    def one_step():
        """walker -> Phi blocks -> Phi^T blocks -> one Phi(Phi^T V).  Returns events + the visit counter only,
        so every step reuses the previous step's device memory (no cudaMalloc in the timed region)."""
        visits = torch.zeros(1, dtype=torch.int64, device=dev)
        st, e_walk = timed(lambda: engine.run_walker(graph, cfg, lo, hi, visits=visits, count_columns=True))
        phi, e_comp = timed(lambda: engine._blocks_from_staging(st, cfg, graph.n_nodes, _lib.SCALE_MUL_RECIP))
        del st
        phi.row_lo = lo
        def transpose_and_prepare():
            # Phi^T blocks + the one-off matvec preparation (row census, workspaces; with several ranks the
            # list of columns shared between row shards) -- all of it inside the timed phase
            phi.build_transpose()
            return phi.plan(f, T_RHS, group=True if world > 1 else None, merged=False)   # f applied per entry

        plan, e_tr = timed(transpose_and_prepare)
        out = torch.empty((hi - lo, T_RHS), dtype=torch.float32, device=dev)
        _, e_mv = timed(lambda: plan(v, out))
        return dict(visits=visits, nnz=phi.nnz, n_rows=phi.n_rows,
                    events=dict(walk=e_walk, compact=e_comp, transpose=e_tr, matvec=e_mv))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                 # nvidia-smi needs ~0.1 s to start: launch it before warm-up
    for _ in range(args.warmup):
        r = one_step()
        del r
    if rank == 0:
        sampler.wait_first_sample()
    # no Python garbage-collector pause inside the timed region (a gen-2 collection on one rank stalls the
    # other ranks in the matvec's all-reduce for about a millisecond)
    gc.collect()
    gc.disable()
    barrier()
    t_wall0 = time.perf_counter()
    results = [one_step() for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    gc.enable()

    phase_ms = {k: sum(r["events"][k][0].elapsed_time(r["events"][k][1]) for r in results)
                for k in ("walk", "compact", "transpose", "matvec")}
    step_ms_total = sum(phase_ms.values())
    if os.environ.get("GRF_BENCH_DEBUG"):
        print(f"[rank {rank}] per-phase ms over {args.steps} steps: "
              + ", ".join(f"{k} {v:.3f}" for k, v in phase_ms.items()), file=sys.stderr, flush=True)
    visits_total = sum(int(r["visits"].item()) for r in results)
    nnz = results[-1]["nnz"]
    n_rows = results[-1]["n_rows"]

    # ---- the matvec as a CG solve uses it: Phi_f merged on the union pattern once per modulator, then
    # many products; every product timed separately with an L2 flush in front
    st = engine.run_walker(graph, cfg, lo, hi)
    phi_cg = engine._blocks_from_staging(st, cfg, graph.n_nodes, _lib.SCALE_MUL_RECIP)
    del st
    phi_cg.row_lo = lo
    phi_cg.shared_hint = shared_hint
    phi_cg.build_transpose()
    _, e_union = timed(lambda: phi_cg.build_union())
    cg_plan, e_mat = timed(lambda: phi_cg.plan(f, T_RHS, group=True if world > 1 else None, merged=True))
    out_cg = torch.empty((hi - lo, T_RHS), dtype=torch.float32, device=dev)
    for _ in range(3):
        cg_plan(v, out_cg)
    cg_events = [timed(lambda: cg_plan(v, out_cg))[1] for _ in range(20)]
    torch.cuda.synchronize(dev)
    cg_ms = sorted(a.elapsed_time(b) for a, b in cg_events)
    cg_info = {"union_build_ms": e_union[0].elapsed_time(e_union[1]), "materialize_ms": e_mat[0].elapsed_time(e_mat[1]),
               "matvec_ms_median": cg_ms[len(cg_ms) // 2], "matvec_ms_min": cg_ms[0], "nnz_union": phi_cg.nnz_union}
    del phi_cg, cg_plan
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the public drop-in call with HOST buffers (host CSR in, scipy CSR list out) + host matvec
    from efficient_graph_gp_sparse.random_walk_samplers_sparse.sparse_sampler import SparseRandomWalk

    del results
    torch.cuda.empty_cache()
    e2e_time, e2e_visits, h2d, d2h = 0.0, 0, 0, 0
    v_host = v.cpu().pin_memory()
    out_host = torch.empty((hi - lo, T_RHS), dtype=torch.float32).pin_memory()   # the product is read back here
    e2e_steps = max(1, min(args.steps, 3))
    for i in range(1 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        rw = SparseRandomWalk(lap, seed=SEED, device=dev)
        steps = rw.get_step_matrices_device(W, P_HALT, L, start_lo=lo, start_hi=hi)
        mats = steps.to_scipy()                          # D2H of every M_l (float64 + int32 + offsets)
        phi_e = engine.PhiBlocks.from_step_matrices(steps)
        phi_e.row_lo = lo
        phi_e.shared_hint = shared_hint
        vd = v_host.to(dev, non_blocking=True)
        out_host.copy_(phi_e.plan(f, T_RHS, group=True if world > 1 else None, merged=False)(vd),
                       non_blocking=True)                # one product, read back into pinned memory
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_time += dt
            e2e_visits += steps.visits
            h2d = lap.indptr.nbytes + lap.indices.nbytes + lap.data.nbytes + v_host.numel() * 4
            d2h = steps.offsets.numel() * 8 + steps.col.numel() * 4 + steps.val.numel() * 8 + out_host.numel() * 4
        del rw, steps, mats, phi_e

    # ---- reduce over ranks: max time, sum of work
    red = torch.tensor([step_ms_total, phase_ms["walk"], phase_ms["compact"], phase_ms["transpose"],
                        phase_ms["matvec"], e2e_time, cg_info["matvec_ms_median"], cg_info["matvec_ms_min"],
                        cg_info["union_build_ms"], cg_info["materialize_ms"]], dtype=torch.float64, device=dev)
    work = torch.tensor([visits_total, e2e_visits, nnz], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    red, work = red.tolist(), work.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        K = args.steps
        step_ms, walk_ms, comp_ms, tr_ms, mv_ms = (x / K for x in red[:5])
        value = work[0] / (red[0] * 1e-3)
        # matvec algorithmic bytes, this rank's launch pair: 2*nnz*8 + 2*L*(rows+1)*4 + 4*N_eff*t*4
        n_cols = graph.n_nodes
        mv_bytes = 2 * nnz * 8 + L * (n_rows + 1) * 4 + L * (n_cols + 1) * 4 + (2 * n_rows + 2 * n_cols) * T_RHS * 4
        mv_gbs = mv_bytes / (mv_ms * 1e-3) / 1e9
        cg_ms_med, cg_ms_min, union_ms, mat_ms = red[6], red[7], red[8], red[9]
        cg_gbs = mv_bytes / (cg_ms_med * 1e-3) / 1e9
        nnz_u = cg_info["nnz_union"]
        layout_bytes = 2 * nnz_u * 8 + (n_rows + 1) * 4 + (n_cols + 1) * 4 + (2 * n_rows + 2 * n_cols) * T_RHS * 4
        traffic = load_traffic()
        walk_bytes = (work[0] / world / K) * WALK_BYTES_PER_STEP
        walk_gbs = walk_bytes / (walk_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 walk loads / f32 matvec", "data": "synthetic",
            "config": workload_config(world, n),
            "phases_ms": {"walk_merge": walk_ms, "compact_blocks": comp_ms, "transpose_blocks": tr_ms,
                          "matvec_phi_phiT_v": mv_ms},
            "phi_build_ms": walk_ms + comp_ms + tr_ms,
            "walker_steps_per_sec": (work[0] / K) / (walk_ms * 1e-3),
            "matvec": {"ms": mv_ms, "algorithmic_gbs": mv_gbs, "nnz_phi_rank0": nnz, "t": T_RHS,
                       "includes_allreduce": world > 1,
                       "what": "per-length Phi blocks, modulator applied per entry (inside the timed step)"},
            "cg_matvec": {"ms": cg_ms_med, "ms_min": cg_ms_min, "algorithmic_gbs": cg_gbs,
                          "layout_gbs": layout_bytes / (cg_ms_med * 1e-3) / 1e9, "nnz_union_rank0": nnz_u,
                          "union_build_ms_once_per_phi": union_ms, "materialize_ms_once_per_modulator": mat_ms,
                          "includes_allreduce": world > 1,
                          "what": "Phi_f merged on the union pattern (MatvecPlan default): 20 products, L2 flushed "
                                  "before each, median"},
            "roofline": {"kernel": "walk_merge_kernel (dominant by time)", "bound": "hbm", "achieved": walk_gbs,
                         "peak": peak, "unit": "GB/s", "frac": walk_gbs / peak,
                         "traffic": traffic.get("walk_merge_bytes"),
                         "peak_source": peak_src,
                         "note": "latency/sector-bound random gathers: 32 algorithmic B per walk-step, "
                                 "graph L2-resident at this size"},
            "roofline_matvec": {"kernel": "spmm_blocks_kernel x2 (Phi^T V, Phi U) on merged Phi_f (CG path)",
                                "bound": "hbm", "achieved": cg_gbs, "peak": peak, "unit": "GB/s",
                                "frac": cg_gbs / peak, "traffic": traffic.get("spmm_merged_pair_bytes"),
                                "peak_source": peak_src, "algorithmic_bytes": mv_bytes,
                                "layout_bytes": layout_bytes,
                                "note": "algorithmic bytes = SURVEY 8d formula on the per-length matrices "
                                        "(2*nnz*8 + 2*L*(n+1)*4 + 4*N*t*4); the merged layout streams fewer "
                                        "(layout_bytes); per-length kernel: see matvec"},
            "e2e": {"value": work[1] / red[5], "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * red[5] / e2e_steps,
                    "what": "SparseRandomWalk(host scipy CSR) -> list of host scipy CSR step matrices, plus one "
                            "host-V -> Phi(Phi^T V) -> host matvec"},
            "gpu_launches": K * 18,
            "gpu_launches_per_step": {"walk_merge": 1, "scan": 3 + 3, "compact_blocks": 1,
                                      "transpose_fill": 1, "transpose_sort": 5, "row_census": 2, "spmm_blocks": 2},
            "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(lap)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(lap):
    """The reference's CPU algorithm on this box's host cores, bounded sample (about 10-30 s)."""
    from oracle import cpu_baseline

    cores = os.cpu_count() or 1
    n = lap.shape[0]
    n_sample = 4 * GRID_NX * cores                        # four grid lines per core: ~1.3e4 walk-steps/core/line
    starts = np.arange(min(n, n_sample))
    dt, visits = cpu_baseline.time_sampler(lap, W, P_HALT, L, starts, n_processes=cores)
    out = {"value": visits / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{len(starts)} of {n} start nodes (x{W} walks, {visits} walk-steps, {dt:.1f} s), pure-Python "
                     f"fork-pool port of sparse_sampler.py:72-132 in oracle/cpu_baseline.py"}
    # the C restatement of the same algorithm, single thread, for scale
    from oracle import c_oracle

    t0 = time.perf_counter()
    _, vis_c = c_oracle.step_matrices(lap, W, P_HALT, L, seed=SEED, start_lo=0, start_hi=min(n, 20 * GRID_NX),
                                      return_visits=True)
    out["c_port_single_thread"] = {"value": vis_c / (time.perf_counter() - t0), "unit": UNIT, "cores": 1}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
