"""BASELINE configs 4 and 5 on one GPU: GraphCNF at the Zinc250k shape (N 38, 9 node types, 3 bond types + none; nodes d=6 K=16,
edges d=2 K=8; flows 4,6,6; hidden 384 / 192; 4 layers) - forward (log-likelihood) at the per-GPU batch 64 of config 4 and
reverse (sampling) at the per-GPU batch 1024 of config 5.   python tools/bench_graphcnf.py [--reps 5] [--profile]"""
import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--train", action="store_true", help="also time a training step (forward in training mode + backward) at --fwd-batch")
ap.add_argument("--cprofile", action="store_true", help="host-side cProfile of the forward pass (where the launch-bound time goes)")
ap.add_argument("--fwd-batch", type=int, default=64)
ap.add_argument("--inv-batch", type=int, default=1024)
ap.add_argument("--gemm-shapes", action="store_true", help="per-shape table of the projection launches of one forward pass (ops.linear_profile)")
args = ap.parse_args()

from categoricalnf_b200 import ops
from categoricalnf_b200.experiments.molecule_generation import GraphCNF

N = 38


class Zinc:
    max_num_nodes = staticmethod(lambda: N)
    num_node_types = staticmethod(lambda: 9)
    num_edge_types = staticmethod(lambda: 3)
    num_max_neighbours = staticmethod(lambda: 4)
    get_node_prior = staticmethod(lambda data_root="data/": np.zeros(9, dtype=np.float32))
    get_edge_prior = staticmethod(lambda data_root="data/": np.zeros(3, dtype=np.float32))


def enc(d):
    return {"use_dequantization": False, "use_variational": False, "use_decoder": False, "num_dimensions": d,
            "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128}, "decoder_config": {"num_layers": 1, "hidden_size": 64}}


params = {"categ_encoding_nodes": enc(6), "categ_encoding_edges": enc(2), "coupling_hidden_size_nodes": 384, "coupling_hidden_size_edges": 192,
          "coupling_num_flows": "4,6,6", "coupling_hidden_layers": 4, "coupling_num_mixtures_nodes": 16, "coupling_num_mixtures_edges": 8,
          "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0}


def molecules(gen, B):
    """Zinc-shaped synthetic graphs: random spanning tree + a few ring-closing edges, degree <= 4, lengths U{20..38}."""
    length = torch.randint(20, N + 1, (B,), generator=gen)
    adj = torch.zeros(B, N, N, dtype=torch.long)
    for b in range(B):
        n = int(length[b])
        deg = [0] * n
        for i in range(1, n):
            cand = [j for j in range(i) if deg[j] < 4]
            j = cand[int(torch.randint(0, len(cand), (1,), generator=gen))]
            t = int(torch.randint(1, 4, (1,), generator=gen))
            adj[b, i, j] = adj[b, j, i] = t
            deg[i] += 1
            deg[j] += 1
        for _ in range(3):
            i, j = (int(v) for v in torch.randint(0, n, (2,), generator=gen))
            if i != j and adj[b, i, j] == 0 and deg[i] < 4 and deg[j] < 4:
                adj[b, i, j] = adj[b, j, i] = 1
                deg[i] += 1
                deg[j] += 1
    x = torch.randint(0, 9, (B, N), generator=gen) * (torch.arange(N)[None, :] < length[:, None]).long()
    return x, adj, length


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
gen = torch.Generator().manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = GraphCNF(params, Zinc).cuda().eval()
x, adj, length = molecules(gen, max(args.fwd_batch, 64))
xc, ac, lc = x.cuda(), adj.cuda(), length.cuda()
with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
    model.initialize_data_dependent([(xc[:64], {"adjacency": ac[:64], "length": lc[:64]})])
Bf, Bi = args.fwd_batch, args.inv_batch
with torch.no_grad():
    fwd = lambda: model(xc[:Bf], adjacency=ac[:Bf], length=lc[:Bf])
    fwd()
    n0 = ops.launch_count()
    fwd()
    launches = ops.launch_count() - n0
    fwd_ms = timed(fwd, args.reps)
    len_i = torch.randint(20, N + 1, (Bi,), generator=gen).cuda()
    z_nodes = model.prior_distribution.sample(shape=(Bi, N, 6)) * (torch.arange(N, device="cuda")[None, :, None] < len_i[:, None, None])
    inv = lambda: model(z_nodes, reverse=True, length=len_i)
    inv_ms = timed(inv, max(2, args.reps // 2))
out = {"config": "GraphCNF Zinc250k shape (N=38, flows 4/6/6, hidden 384/192, 4 layers)", "fwd_batch": Bf, "fwd_ms": fwd_ms,
       "fwd_graphs_per_s": Bf / fwd_ms * 1e3, "cnf_launches_per_forward": launches, "sampling_batch": Bi, "sampling_ms": inv_ms,
       "sampling_graphs_per_s": Bi / inv_ms * 1e3}
if True:
    from categoricalnf_b200.experiments.molecule_generation import GraphedLogLikelihood
    graphed = GraphedLogLikelihood(model)
    gfwd = lambda: graphed(xc[:Bf], ac[:Bf], lc[:Bf])
    with torch.no_grad():
        gfwd()
        out["graphed_fwd_ms"] = timed(gfwd, args.reps * 2)
    out["graphed_fwd_graphs_per_s"] = Bf / out["graphed_fwd_ms"] * 1e3
    out["graph_captures"] = graphed.captures
if args.gemm_shapes:
    ops.linear_profile = []
    with torch.no_grad():
        fwd()
    torch.cuda.synchronize()
    prof, ops.linear_profile = ops.linear_profile, None
    tab = {}
    for (M, Nn, K, prec, e0, e1) in prof:
        t = tab.setdefault((M, Nn, K, prec), [0, 0.0])
        t[0] += 1
        t[1] += e0.elapsed_time(e1)
    rows = sorted(tab.items(), key=lambda kv: -kv[1][1])
    out["gemm_shapes"] = [{"M": k[0], "N": k[1], "K": k[2], "precision": k[3], "launches": v[0], "total_ms": round(v[1], 3),
                           "mean_us": round(v[1] / v[0] * 1e3, 1),
                           "mma_tflops": round((3 if k[3] == "3xtf32" else 1) * 2.0 * k[0] * k[1] * k[2] * v[0] / v[1] / 1e9, 1)} for k, v in rows]
    out["gemm_total_ms"] = sum(v[1] for v in tab.values())
if args.train:
    model.train()
    params_ = [p_ for p_ in model.parameters() if p_.requires_grad]

    def step():
        for p_ in params_:
            p_.grad = None
        zt, lt = model(xc[:Bf], adjacency=ac[:Bf], length=lc[:Bf])
        (-(lt.mean())).backward()
    step()
    n0 = ops.launch_count()
    step()
    out["cnf_launches_per_train_step"] = ops.launch_count() - n0
    out["train_step_ms"] = timed(step, max(2, args.reps // 2))
    out["train_graphs_per_s"] = Bf / out["train_step_ms"] * 1e3
    from categoricalnf_b200.experiments.molecule_generation import GraphedTrainingStep
    gstep_ = GraphedTrainingStep(model)
    gstep = lambda: gstep_(xc[:Bf], ac[:Bf], lc[:Bf])
    gstep()
    out["graphed_train_step_ms"] = timed(gstep, max(3, args.reps))
    out["graphed_train_graphs_per_s"] = Bf / out["graphed_train_step_ms"] * 1e3
    model.eval()
print(json.dumps(out))
if args.cprofile:
    import cProfile
    import pstats
    pr = cProfile.Profile()
    with torch.no_grad():
        pr.enable()
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)
    st.sort_stats("cumulative").print_stats(35)
if args.profile:
    from torch.profiler import profile, ProfilerActivity
    for tag, fn in (("forward B=%d" % Bf, fwd), ("sampling B=%d" % Bi, inv)):
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            fn()
            torch.cuda.synchronize()
        print("==", tag)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
