"""BASELINE config 2 (316 x 316 grid) per-phase device times on one GPU -- bench.py's config2 section on its own.
  [GRF_B200_WALK_WIDE=0|1] python profiles/prof_cfg2.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "efficient-gaussian-process-on-graphs_b200")]
import torch
import bench
from grf_b200 import _lib, engine

torch.cuda.set_device(0)
res = bench.config2_leg(torch, engine, _lib, torch.device("cuda", 0), 0)
print(json.dumps({k: res[k] for k in ("value", "ms_per_step", "phases_ms", "walker_steps_per_sec", "matvec_per_length",
                                      "cg_matvec_merged")}))
