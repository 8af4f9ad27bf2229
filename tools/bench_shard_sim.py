"""Exercise bench_graphs at the PER-RANK shapes of an 8-GPU run on one GPU (global batch 512 -> 64 per rank: the
CUDA-graph replay path of config 4; 8192 -> 1024 per rank for config 5) - a dry run of what torchrun N = 8 executes per rank.
    python tools/bench_shard_sim.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_graphs
import graph_workloads as G

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=8)
a = ap.parse_args()
G.MOL["B_fwd"] //= a.ranks
G.MOL["B_inv"] //= a.ranks
args = argparse.Namespace(steps=5, graph_steps=5, no_cpu=True, no_train=False)
torch.cuda.set_device(0)
recs = bench_graphs.run_all(args, 0, 1, torch.device("cuda", 0), None)
for r in recs:
    print(json.dumps({k: r.get(k) for k in ("name", "value", "ms_per_step", "batch_per_gpu", "mode", "error", "traceback")})[:600])
