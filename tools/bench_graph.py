"""BASELINE config 3 (graph colouring, B 1024, N 20, GraphNodeFlow defaults: 8 flows, hidden 384, 4 attention layers,
8 mixtures, d=2): forward (log-likelihood) and reverse (sampling) throughput of the drop-in flow on one GPU, with a
kernel breakdown.   python tools/bench_graph.py [--batch 1024] [--reps 10] [--profile] [--cpu]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--nodes", type=int, default=20)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on a 32-graph sample")
ap.add_argument("--train", action="store_true", help="also time a training step (forward in training mode + backward)")
ap.add_argument("--config", default="gc", choices=["gc", "mol"],
                help="gc: graph colouring (config 3); mol: GraphCNF molecule generation forward (config 4, batch 64 per GPU) and "
                     "sampling (config 5, batch 1024 per GPU)")
args = ap.parse_args()
if args.config == "mol":
    import runpy
    sys.argv = [sys.argv[0]] + (["--profile"] if args.profile else []) + ["--reps", str(args.reps)]
    runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_graphcnf.py"), run_name="__main__")
    sys.exit(0)

from categoricalnf_b200 import ops
from categoricalnf_b200.experiments.graph_coloring import GraphNodeFlow
from categoricalnf_b200.experiments.graph_coloring.graph_node_flow import length_masks


class Dataset:
    @staticmethod
    def num_node_types():
        return 3


def params():
    return {"categ_encoding": {"use_dequantization": False, "use_variational": False, "use_decoder": False, "num_dimensions": 2,
                               "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128},
                               "decoder_config": {"num_layers": 1, "hidden_size": 64}},
            "coupling_num_flows": 8, "coupling_hidden_size": 384, "coupling_hidden_layers": 4, "coupling_num_mixtures": 8,
            "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0}


def graphs(gen, B, N, p=0.15):
    length = torch.randint(11, N, (B,), generator=gen)
    up = torch.triu((torch.rand(B, N, N, generator=gen) < p).long(), diagonal=1)
    adj = up + up.transpose(1, 2)
    valid = torch.arange(N)[None, :] < length[:, None]
    return adj * (valid[:, :, None] & valid[:, None, :]).long(), length


torch.manual_seed(0)
gen = torch.Generator().manual_seed(0)
B, N = args.batch, args.nodes
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    model = GraphNodeFlow(params(), Dataset).cuda().eval()
adj, length = graphs(gen, B, N)
x = torch.randint(0, 3, (B, N), generator=gen)
xc, ac, lc = x.cuda(), adj.cuda(), length.cuda()
with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
    model.initialize_data_dependent([(xc[:256], {"adjacency": ac[:256], "length": lc[:256]})])


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    z, ldj = model(xc, adjacency=ac, length=lc)
    n0 = ops.launch_count()
    model(xc, adjacency=ac, length=lc)
    launches = ops.launch_count() - n0
    fwd_ms = timed(lambda: model(xc, adjacency=ac, length=lc), args.reps)
    kpm, cpm = length_masks(lc, N)
    kw = dict(adjacency=ac, length=lc, channel_padding_mask=cpm, src_key_padding_mask=kpm)

    def reverse():
        zz = z
        for layer in reversed(list(model.flow_layers)):
            zz = layer(zz, reverse=True, **kw)[0]
        return zz
    rev_ms = timed(reverse, args.reps)
out = {"config": "graph_coloring B=%d N=%d (8 flows, hidden 384, 4 attention layers, K=8, d=2)" % (B, N), "fwd_ms": fwd_ms,
       "fwd_graphs_per_s": B / fwd_ms * 1e3, "reverse_ms": rev_ms, "reverse_graphs_per_s": B / rev_ms * 1e3,
       "cnf_launches_per_forward": launches}
if args.train:
    model.train()
    params_ = [p_ for p_ in model.parameters() if p_.requires_grad]

    def step():
        for p_ in params_:
            p_.grad = None
        zt, lt = model(xc, adjacency=ac, length=lc)
        (-(lt.sum()) + 0.5 * (zt ** 2).sum()).backward()
    n0 = ops.launch_count()
    step()
    out["cnf_launches_per_train_step"] = ops.launch_count() - n0
    out["train_step_ms"] = timed(step, max(3, args.reps // 2))
    out["train_graphs_per_s"] = B / out["train_step_ms"] * 1e3
    model.eval()
if args.cpu:
    from oracle import graph_oracle as GO
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    nb = 32
    u = torch.rand(nb * N, 1, 2, generator=gen)
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    GO.graph_node_flow(sd, x[:nb], adj[:nb], length[:nb], u, num_flows=8, num_layers=4, num_mixtures=8)
    dt = time.perf_counter() - t0
    out["cpu_oracle_graphs_per_s"] = nb / dt
    out["cpu_cores"] = os.cpu_count()
print(json.dumps(out))
if args.profile and args.train:
    from torch.profiler import profile, ProfilerActivity
    model.train()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
elif args.profile:
    from torch.profiler import profile, ProfilerActivity
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            model(xc, adjacency=ac, length=lc)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
