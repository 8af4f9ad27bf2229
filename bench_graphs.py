"""BASELINE.json configs 3, 4, 5 as driver-measured records of bench.py's JSON line (key ``configs``).

    graph_coloring        config 3: GraphNodeFlow log-likelihood pass, batch 1024 PER GPU (weak scaling, like the headline)
    molecule_generation   config 4: GraphCNF log-likelihood pass, GLOBAL batch 512 sharded over the ranks (strong scaling)
    inverse_sampling      config 5: GraphCNF reverse (sampling) pass, GLOBAL batch 8192 sharded over the ranks (strong scaling)

One step = one pass of the model over the rank's shard; the ranks all-reduce (sum log-likelihood, count) once per step on
the log-likelihood configs (sampling has no exchange).  Every record carries
  value / ms_per_step   device-timed (CUDA events, barrier on both sides, max over ranks), inputs resident in HBM
  e2e                   the same pass driven from pinned HOST inputs: H2D of the graphs, D2H of the per-graph
                        log-likelihoods / the sampled graphs inside the timed region
  roofline              tensor-pipe roofline of the dominant kernel, the 3xTF32 tcgen05 projection GEMM (cnf_linear_fwd): CUDA
                        events around every projection launch, MMA flops (3 passes x 2MNK) / time against the TF32 peak =
                        MEASURED_PEAKS.json bf16_tflops_sustained / 2
  cpu_baseline          the unmodified reference model (baseline/_ref, state dict of the GPU model loaded by name) on the host
                        cores, on a bounded sub-batch (rank 0, N = 1 only)
  parity                the GPU model against that reference model on identical graphs and recorded noise
"""
from __future__ import annotations

import json
import os
import time

import torch

import graph_workloads as G
import workload as W

ROOT = os.path.dirname(os.path.abspath(__file__))


def tensor_peak():
    """TF32 dense peak in TFLOP/s: half the measured sustained bf16 cuBLAS rate (tcgen05 kind::tf32 runs at half the bf16
    rate); fallback = half of B200_PROFILING.md's nominal 2250 bf16."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["bf16_tflops_sustained"]) / 2, "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 = half the bf16 rate)"
    return 1125.0, "fallback: B200_PROFILING.md nominal 2250 TFLOP/s bf16 / 2"


class Ctx:
    def __init__(self, args, rank, world, dev, dist):
        self.args, self.rank, self.world, self.dev, self.dist = args, rank, world, dev, dist
        self.distributed = world > 1
        self.steps = max(2, min(args.steps, args.graph_steps))
        self.warmup = 3
        from categoricalnf_b200.sharding import LogLikAllReducer
        self.reducer = LogLikAllReducer(dev, slots=4)

    def barrier(self):
        self.reducer.finish()
        if self.distributed:
            self.dist.barrier()
        torch.cuda.synchronize()

    def loglik(self, ops, z, ldj, pad):
        """Per-graph log-likelihood ldj + log p(z) and the step's (sum, count) pair from ONE kernel (cnf_logistic_logprob
        add / total); the pair is all-reduced over the ranks on the communication stream."""
        ll, _ = ops.logistic_logprob(z, pad=pad, add=ldj, total=self.reducer.slot())
        self.reducer.reduce()
        return ll

    def timed(self, fn, steps=None):
        """ms per step of ``fn`` (max over ranks)."""
        steps = steps or self.steps
        for _ in range(self.warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        self.reducer.finish()
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def gemm_roofline(ops, fn, step_ms, note):
    """Run ``fn`` once with CUDA events around every cnf_linear_fwd launch; the dominant GEMM shape (largest summed time)
    against the TF32 tensor peak."""
    torch.cuda.synchronize()
    ops.linear_profile = []
    try:
        fn()
        torch.cuda.synchronize()
        prof = ops.linear_profile
    finally:
        ops.linear_profile = None
    if not prof:
        return None
    shapes, total_ms, total_flops = {}, 0.0, 0.0
    for M, N, K, precision, e0, e1 in prof:
        ms = e0.elapsed_time(e1)
        passes = 3 if precision == "3xtf32" else 1
        s = shapes.setdefault((M, N, K, precision), [0.0, 0.0, 0])
        s[0] += ms
        s[1] += 2.0 * M * N * K * passes
        s[2] += 1
        total_ms += ms
        total_flops += 2.0 * M * N * K * passes
    (M, N, K, precision), (ms, flops, count) = max(shapes.items(), key=lambda kv: kv[1][0])
    peak, src = tensor_peak()
    achieved = flops / (ms * 1e-3) / 1e12
    passes = 3 if precision == "3xtf32" else 1
    return {"kernel": "linear_tc_kernel<%s> (cnf_linear_fwd), M=%d N=%d K=%d" % (precision, M, N, K), "bound": "tensor",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
            "flops": "MMA flops: %d pass(es) x 2MNK per launch (useful fp32-accurate rate = achieved / %d = %.1f TFLOP/s)"
                     % (passes, passes, achieved / passes),
            "peak_source": src, "mean_launch_ms": ms / count, "launches_timed": count,
            "all_projections": {"launches": len(prof), "ms_per_step": total_ms, "mma_tflops": total_flops / (total_ms * 1e-3) / 1e12,
                                "share_of_step": total_ms / step_ms},
            "timed": note}


def _dev_over_tol(a, b, rtol, atol):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / (rtol * b.abs() + atol)).max())


def _cpu_time(fn, n_graphs, budget_s=8.0, max_reps=5):
    torch.set_num_threads(os.cpu_count() or 1)
    fn()                    # warm-up at the measured size
    times, t_start = [], time.perf_counter()
    while len(times) < 2 or (time.perf_counter() - t_start < budget_s and len(times) < max_reps):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return n_graphs * len(times) / sum(times), len(times), sum(times)


# --------------------------------------------------------------------------------------------------------------------
def run_graph_coloring(c: Ctx):
    from categoricalnf_b200 import ops
    dev, B, N = c.dev, G.GC["B"], G.GC["N"]
    model = G.build_gc_model(dev, seed=0)
    gen = torch.Generator().manual_seed(100 + c.rank)
    x, adj, length = G.gc_graphs(gen, B)
    xc, ac, lc = x.to(dev), adj.to(dev), length.to(dev)
    x0, a0, l0 = G.gc_graphs(torch.Generator().manual_seed(7), G.GC["init_batch"])        # same init batch on every rank
    G.data_init(model, x0.to(dev), a0.to(dev), l0.to(dev))
    pad = (torch.arange(N, device=dev)[None, :] < lc[:, None]).float()

    def step(xi=xc, ai=ac, li=lc):
        with torch.no_grad():
            z, ldj = model(xi, adjacency=ai, length=li)
            ll = c.loglik(ops, z, ldj, pad)
        return ll

    step()
    n0 = ops.launch_count()
    step()
    launches = ops.launch_count() - n0
    ms = c.timed(step)
    roof = gemm_roofline(ops, step, ms, "one extra pass right after the timed region, events around every projection launch")
    ops.check_status(dev, "bench graph_coloring")

    hx, ha, hl = x.pin_memory(), adj.pin_memory(), length.pin_memory()
    h_ll = torch.empty(B, dtype=torch.float32).pin_memory()
    dx, da, dl = torch.empty_like(xc), torch.empty_like(ac), torch.empty_like(lc)

    def e2e_step():
        dx.copy_(hx, non_blocking=True)
        da.copy_(ha, non_blocking=True)
        dl.copy_(hl, non_blocking=True)
        ll = step(dx, da, dl)
        h_ll.copy_(ll, non_blocking=True)
        torch.cuda.current_stream().synchronize()         # the host consumes this step's log-likelihoods

    e2e_ms = c.timed(e2e_step)
    rec = {"name": "graph_coloring", "baseline_config": "configs[2]", "metric": "GraphNodeFlow fwd+ldj graphs/sec", "unit": "graphs/s",
           "value": B * c.world / (ms * 1e-3), "ms_per_step": ms, "steps": c.steps, "warmup": c.warmup, "scaling": "weak",
           "batch_per_gpu": B, "global_batch": B * c.world, "nodes": N, "mode": "eager, %d C-ABI launches per pass" % launches,
           "gpu_launches_per_step": launches, "dtype": "f32 (projections 3xTF32)",
           "workload": "synthetic 3-colour graphs N=20, GraphNodeFlow defaults (8 flows, hidden 384, 4 RGCN attention layers, K=8, d=2)",
           "e2e": {"value": B * c.world / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": B * N * 8 + B * N * N * 8 + B * 8, "d2h_bytes_per_step": B * 4},
           "roofline": roof}
    if c.rank == 0 and not c.args.no_cpu:
        rec.update(_gc_reference_legs(c, model, ops))
    return rec


def _gc_reference_legs(c, model, ops):
    out = {}
    try:
        ref = G.build_reference_like(model, "gc")
    except Exception as exc:      # noqa: BLE001 - baseline/_ref did not travel: say so instead of failing the bench
        return {"cpu_baseline": {"unavailable": "%s: %s" % (type(exc).__name__, exc)}, "parity": None}
    dev, N = c.dev, G.GC["N"]
    nb = 8
    x, adj, length = G.gc_graphs(torch.Generator().manual_seed(31), nb)
    rec = G.NoiseRecorder(G.reference_encodings(ref, "gc"), seed=5)
    with torch.no_grad():
        z_ref, ldj_ref = ref(x, adjacency=adj, length=length)
        z, ldj = model(x.to(dev), adjacency=adj.to(dev), length=length.to(dev), u_noise=rec.draws[0].to(dev))
    out["parity"] = {"against": "unmodified reference GraphNodeFlow (baseline/_ref), same parameters, graphs and noise", "batch": nb,
                     "z_dev_over_tol": _dev_over_tol(z, z_ref, 1e-4, 2e-5), "ldj_dev_over_tol": _dev_over_tol(ldj, ldj_ref, 1e-4, 2e-4),
                     "z_max_abs": float((z.cpu().double() - z_ref.double()).abs().max()),
                     "ldj_max_rel": float(((ldj.cpu().double() - ldj_ref.double()).abs() / ldj_ref.double().abs()).max()),
                     "tolerance": "|a-b| <= 1e-4|b| + 2e-5 (z through 8 couplings x 4-layer fp32 networks of hidden 384), "
                                  "1e-4|b| + 2e-4 (ldj); <= 1 is inside; ldj_max_rel is the pure relative deviation"}
    if c.world == 1:
        nb = 256
        x, adj, length = G.gc_graphs(torch.Generator().manual_seed(32), nb)

        def cpu():
            with torch.no_grad():
                ref(x, adjacency=adj, length=length)
        value, reps, total = _cpu_time(cpu, nb)
        out["cpu_baseline"] = {"value": value, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "reference",
                               "sample": "%d passes over %d of the 1024 graphs, unmodified reference GraphNodeFlow from baseline/_ref, "
                                         "%.1f s of CPU work" % (reps, nb, total)}
    return out


# --------------------------------------------------------------------------------------------------------------------
def run_molecules(c: Ctx):
    """configs 4 and 5 share the model: one build + data-dependent init, two records."""
    from categoricalnf_b200 import ops
    from categoricalnf_b200.experiments.molecule_generation import GraphedLogLikelihood
    dev, N = c.dev, G.MOL["N"]
    model = G.build_mol_model(dev, seed=0)
    x0, a0, l0 = G.molecules(torch.Generator().manual_seed(7), G.MOL["init_batch"])
    G.data_init(model, x0.to(dev), a0.to(dev), l0.to(dev))
    records = []

    # ---- config 4: log-likelihood, global batch 512 -------------------------------------------------------------------
    Bg = G.MOL["B_fwd"]
    lo, hi = _shard(Bg, c.rank, c.world)
    B = hi - lo
    x, adj, length = G.molecules(torch.Generator().manual_seed(200), Bg)
    x, adj, length = x[lo:hi].contiguous(), adj[lo:hi].contiguous(), length[lo:hi].contiguous()
    xc, ac, lc = x.to(dev), adj.to(dev), length.to(dev)
    graphed = GraphedLogLikelihood(model) if B <= G.MOL.get("graph_replay_max", 512) else None

    def fwd(xi=xc, ai=ac, li=lc):
        with torch.no_grad():
            if graphed is not None:
                z, ldj = graphed(xi, ai, li)
            else:
                z, ldj = model(xi, adjacency=ai, length=li)
            pad = (torch.arange(N, device=dev)[None, :] < li[:, None]).float()
            ll = c.loglik(ops, z, ldj, pad)
        return ll

    def fwd_eager():
        with torch.no_grad():
            model(xc, adjacency=ac, length=lc)

    fwd()
    n0 = ops.launch_count()
    fwd_eager()
    launches = ops.launch_count() - n0
    ms = c.timed(fwd)
    roof = gemm_roofline(ops, fwd_eager, ms, "one extra eager pass right after the timed region (events cannot be recorded inside a "
                                             "CUDA-graph replay), events around every projection launch")
    ops.check_status(dev, "bench molecule_generation")
    hx, ha, hl = x.pin_memory(), adj.pin_memory(), length.pin_memory()
    h_ll = torch.empty(B, dtype=torch.float32).pin_memory()
    dx, da, dl = torch.empty_like(xc), torch.empty_like(ac), torch.empty_like(lc)

    def e2e_fwd():
        dx.copy_(hx, non_blocking=True)
        da.copy_(ha, non_blocking=True)
        dl.copy_(hl, non_blocking=True)
        ll = fwd(dx, da, dl)
        h_ll.copy_(ll, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_ms = c.timed(e2e_fwd)
    rec = {"name": "molecule_generation", "baseline_config": "configs[3]", "metric": "GraphCNF fwd+ldj graphs/sec", "unit": "graphs/s",
           "value": Bg / (ms * 1e-3), "ms_per_step": ms, "steps": c.steps, "warmup": c.warmup, "scaling": "strong",
           "global_batch": Bg, "batch_per_gpu": B, "nodes": N,
           "mode": ("whole pass replayed from a CUDA graph (GraphedLogLikelihood, %d kernels)" % launches) if graphed is not None
           else "eager, %d C-ABI launches per pass" % launches,
           "gpu_launches_per_step": launches, "dtype": "f32 (projections 3xTF32)",
           "workload": "Zinc250k-shaped synthetic molecules (N=38, 9 node types, 3 bond types + none), GraphCNF 3-step node / edge / "
                       "adjacency flow (flows 4,6,6; hidden 384/192; 4 layers; K 16/8), %.1f M parameters"
                       % (sum(p.numel() for p in model.parameters()) / 1e6),
           "e2e": {"value": Bg / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": B * N * 8 + B * N * N * 8 + B * 8, "d2h_bytes_per_step": B * 4},
           "roofline": roof}
    records.append(rec)

    # ---- config 5: sampling, global batch 8192 -------------------------------------------------------------------------
    Bg = G.MOL["B_inv"]
    lo, hi = _shard(Bg, c.rank, c.world)
    B = hi - lo
    chunk = min(B, G.MOL["inv_chunk"])
    len_all = torch.randint(20, N + 1, (Bg,), generator=torch.Generator().manual_seed(300))[lo:hi].contiguous()
    len_dev = len_all.to(dev)
    D = G.MOL["D_nodes"]
    ar = torch.arange(N, device=dev)
    out_x = torch.empty(B, N, dtype=torch.int64, device=dev)
    out_a = torch.empty(B, N, N, dtype=torch.int64, device=dev)
    out_l = torch.empty(B, dtype=torch.float32, device=dev)

    def sample(li=len_dev):
        with torch.no_grad():
            for s in range(0, B, chunk):
                l = li[s:s + chunk]
                z_nodes = model.prior_distribution.sample(shape=(l.numel(), N, D)) * (ar[None, :, None] < l[:, None, None])
                (xs, adjs), ldj = model(z_nodes, reverse=True, length=l)
                out_x[s:s + chunk], out_a[s:s + chunk], out_l[s:s + chunk] = xs, adjs, ldj

    steps5 = max(2, min(c.steps, 3))
    sample()
    n0 = ops.launch_count()
    sample()
    launches = ops.launch_count() - n0
    ms = c.timed(sample, steps5)
    roof = gemm_roofline(ops, sample, ms, "one extra pass right after the timed region, events around every projection launch")
    ops.check_status(dev, "bench inverse_sampling")
    h_len = len_all.pin_memory()
    d_len = torch.empty_like(len_dev)
    h_x, h_a = torch.empty(B, N, dtype=torch.int64).pin_memory(), torch.empty(B, N, N, dtype=torch.int64).pin_memory()
    h_l = torch.empty(B, dtype=torch.float32).pin_memory()

    def e2e_sample():
        d_len.copy_(h_len, non_blocking=True)
        sample(d_len)
        h_x.copy_(out_x, non_blocking=True)
        h_a.copy_(out_a, non_blocking=True)
        h_l.copy_(out_l, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_ms = c.timed(e2e_sample, steps5)
    rec5 = {"name": "inverse_sampling", "baseline_config": "configs[4]", "metric": "GraphCNF reverse-pass (sampling) graphs/sec",
            "unit": "graphs/s", "value": Bg / (ms * 1e-3), "ms_per_step": ms, "steps": steps5, "warmup": c.warmup, "scaling": "strong",
            "global_batch": Bg, "batch_per_gpu": B, "nodes": N,
            "mode": "eager, %d chunk(s) of %d graphs per step, %d C-ABI launches per step" % (-(-B // chunk), chunk, launches),
            "gpu_launches_per_step": launches, "dtype": "f32 (projections 3xTF32)",
            "workload": "GraphCNF reverse pass: prior sample -> 3 inverse flow steps (safeguarded-Newton inverse of the mixture CDF "
                        "where the reference bisects) -> decoded node types + adjacency",
            "e2e": {"value": Bg / (e2e_ms * 1e-3), "unit": "graphs/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": B * 8,
                    "d2h_bytes_per_step": B * N * 8 + B * N * N * 8 + B * 4},
            "roofline": roof}
    records.append(rec5)
    if c.rank == 0 and not c.args.no_cpu:
        legs4, legs5 = _mol_reference_legs(c, model)
        rec.update(legs4)
        rec5.update(legs5)
    if not c.args.no_train:
        records.append(run_molecule_training(c, model, ops))
    return records


def run_molecule_training(c: Ctx, model, ops):
    """Training step of config 4 (general/train.py:148-160 on experiments/molecule_generation/task.py's loss): GLOBAL batch
    512 sharded over the ranks; forward in training mode, loss = mean over the shard of -(ldj + log p(z)) / length, backward
    through the backward kernels, gradients all-reduced over NCCL in flat buckets WHILE backward runs
    (sharding.GradientReducer: per-parameter post-accumulate hooks, communication stream), fused Adam step."""
    from categoricalnf_b200.sharding import GradientReducer
    dev, N = c.dev, G.MOL["N"]
    Bg = G.MOL["B_fwd"]
    lo, hi = _shard(Bg, c.rank, c.world)
    B = hi - lo
    x, adj, length = G.molecules(torch.Generator().manual_seed(200), Bg)
    xc, ac, lc = x[lo:hi].contiguous().to(dev), adj[lo:hi].contiguous().to(dev), length[lo:hi].contiguous().to(dev)
    pad3 = (torch.arange(N, device=dev)[None, :] < lc[:, None]).float().unsqueeze(-1)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    n_params = sum(p.numel() for p in params)
    use_graph = B <= G.MOL.get("graph_replay_max_train", 512)
    red = GradientReducer(params, bucket_bytes=32 << 20, profile=True, hooks=not use_graph)
    opt = torch.optim.Adam(params, lr=1e-5, fused=True)
    state = {}

    def train_step():
        red.zero_grad()
        z, ldj = model(xc, adjacency=ac, length=lc)
        logp = (model.prior_distribution.log_prob(z) * pad3).sum(dim=[1, 2])
        loss = (-(ldj + logp) / lc.to(ldj.dtype)).mean()
        loss.backward()
        red.finish()
        opt.step()
        state["loss"] = loss.detach()

    # small shards (N = 8: 64 molecules per GPU) are launch-bound (~2300 launches per step): replay forward + backward from
    # a CUDA graph (GraphedTrainingStep) and reduce the gradients right after the replay
    graphed = None
    if use_graph:
        from categoricalnf_b200.experiments.molecule_generation import GraphedTrainingStep
        ar = torch.arange(N, device=dev)

        def loss_fn(z, ldj, length):
            pad = (ar[None, :] < length[:, None]).to(z.dtype).unsqueeze(-1)
            logp = (model.prior_distribution.log_prob(z) * pad).sum(dim=[1, 2])
            return (-(ldj + logp) / length.to(ldj.dtype)).mean()
        graphed = GraphedTrainingStep(model, loss_fn=loss_fn)

        def train_step():      # noqa: F811
            state["loss"] = graphed(xc, ac, lc)
            red.reduce_now()
            red.finish()
            opt.step()

    steps = max(2, min(c.steps, 3))
    try:
        train_step()
        n0 = ops.launch_count()
        train_step()
        launches = ops.launch_count() - n0
        torch.cuda.synchronize()
        red.comm_stats()
        ms = c.timed(train_step, steps)
        nbytes, comm_ms, bus = red.comm_stats()
        nsteps = steps + c.warmup
        ops.check_status(dev, "bench molecule_generation training")
        mode = ("forward + backward replayed from a CUDA graph (GraphedTrainingStep), gradients copied into %d flat buckets and "
                "all-reduced on a communication stream right after the replay, fused Adam" % len(red.buckets)) if graphed is not None \
            else ("eager; gradients reduced bucket by bucket on a communication stream while backward runs "
                  "(%d flat buckets of <= 32 MiB, p.grad = views), fused Adam" % len(red.buckets))
        rec = {"name": "molecule_generation_train", "baseline_config": "configs[3], training step", "unit": "graphs/s",
               "metric": "GraphCNF training step (fwd + bwd + gradient all-reduce + Adam) graphs/sec",
               "value": Bg / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": c.warmup, "scaling": "strong",
               "global_batch": Bg, "batch_per_gpu": B, "parameters": n_params, "gpu_launches_per_step": launches,
               "mode": mode, "loss": float(state["loss"]), "dtype": "f32 (projections 3xTF32)",
               "collective": {"kind": "NCCL all-reduce (sum) of the flat gradient buckets" +
                                      (", overlapped with backward" if graphed is None else ", after the graph replay"),
                              "bytes_per_step": nbytes / nsteps if c.world > 1 else 0,
                              "comm_stream_ms_per_step": comm_ms / nsteps if c.world > 1 else 0.0,
                              "bus_GBps": bus, "bus_formula": "2 (N-1) / N x bytes / time on the communication stream"}}
    finally:
        red.close()
        model.eval()
        for p in params:
            p.grad = None
    return rec


def _shard(n, rank, world):
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _mol_reference_legs(c, model):
    try:
        ref = G.build_reference_like(model, "mol")
    except Exception as exc:      # noqa: BLE001
        na = {"cpu_baseline": {"unavailable": "%s: %s" % (type(exc).__name__, exc)}, "parity": None}
        return na, dict(na)
    dev, N = c.dev, G.MOL["N"]
    P = N * (N - 1) // 2
    out4, out5 = {}, {}
    nb = 4
    x, adj, length = G.molecules(torch.Generator().manual_seed(41), nb)
    rec = G.NoiseRecorder(G.reference_encodings(ref, "mol"), seed=6)
    with torch.no_grad():
        z_ref, ldj_ref = ref(x, adjacency=adj, length=length)
        u_nodes, u_edges, u_virtual = (d.to(dev) for d in rec.draws[:3])
        z, ldj = model(x.to(dev), adjacency=adj.to(dev), length=length.to(dev), u_noise=u_nodes, u_noise_edges=u_edges,
                       u_noise_virtual=u_virtual)
    out4["parity"] = {"against": "unmodified reference GraphCNF (baseline/_ref), same parameters, molecules and noise", "batch": nb,
                      "z_dev_over_tol": _dev_over_tol(z, z_ref, 1e-4, 2e-5), "ldj_dev_over_tol": _dev_over_tol(ldj, ldj_ref, 1e-4, 5e-4),
                      "z_max_abs": float((z.cpu().double() - z_ref.double()).abs().max()),
                      "ldj_max_rel": float(((ldj.cpu().double() - ldj_ref.double()).abs() / ldj_ref.double().abs()).max()),
                      "tolerance": "|a-b| <= 1e-4|b| + 2e-5 (z through 16 couplings x 4-layer fp32 networks), 1e-4|b| + 5e-4 (ldj, "
                                   "|ldj| ~ 1e3); <= 1 is inside; ldj_max_rel is the pure relative deviation"}
    # sampling parity: same node / edge latents through both reverse passes
    g = torch.Generator().manual_seed(43)
    ns = 2
    len_s = torch.randint(20, N + 1, (ns,), generator=g)
    z_nodes = torch.randn(ns, N, G.MOL["D_nodes"], generator=g) * (torch.arange(N)[None, :, None] < len_s[:, None, None])
    z_edges = torch.randn(ns, P, G.MOL["D_edges"], generator=g)
    ref.prior_distribution.sample = lambda shape=None, temp=1.0, **kw: z_edges
    t0 = time.perf_counter()
    with torch.no_grad():
        (x_ref, a_ref), l_ref = ref(z_nodes, reverse=True, length=len_s)
    t_ref = time.perf_counter() - t0
    with torch.no_grad():
        (x_gpu, a_gpu), l_gpu = model(z_nodes.to(dev), reverse=True, length=len_s.to(dev), z_edges_init=z_edges.to(dev))
    valid = torch.arange(N)[None, :] < len_s[:, None]
    out5["parity"] = {"against": "unmodified reference GraphCNF reverse pass (bisection inverse), same latents", "batch": ns,
                      "adjacency_equal_frac": float((a_gpu.cpu() == a_ref).float().mean()),
                      "node_types_equal_frac": float((x_gpu.cpu()[valid] == x_ref[valid]).float().mean()),
                      "ldj_dev_over_tol": _dev_over_tol(l_gpu, l_ref, 1e-4, 5e-4)}
    if c.world == 1:
        nb = 64
        x, adj, length = G.molecules(torch.Generator().manual_seed(42), nb)

        def cpu():
            with torch.no_grad():
                ref(x, adjacency=adj, length=length)
        value, reps, total = _cpu_time(cpu, nb)
        out4["cpu_baseline"] = {"value": value, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "reference",
                                "sample": "%d passes over %d of the 512 molecules, unmodified reference GraphCNF from baseline/_ref "
                                          "(two torch-2.x compatibility patches at import, graph_workloads.py), %.1f s of CPU work"
                                          % (reps, nb, total)}
        nsb = 32
        g = torch.Generator().manual_seed(44)
        len_b = torch.randint(20, N + 1, (nsb,), generator=g)
        zn_b = torch.randn(nsb, N, G.MOL["D_nodes"], generator=g) * (torch.arange(N)[None, :, None] < len_b[:, None, None])
        ze_b = torch.randn(nsb, P, G.MOL["D_edges"], generator=g)
        ref.prior_distribution.sample = lambda shape=None, temp=1.0, **kw: ze_b

        def cpu_inv():
            with torch.no_grad():
                ref(zn_b, reverse=True, length=len_b)
        value, reps, total = _cpu_time(cpu_inv, nsb)
        out5["cpu_baseline"] = {"value": value, "unit": "graphs/s", "cores": torch.get_num_threads(), "kind": "reference",
                                "sample": "%d reverse passes over %d of the 8192 graphs, unmodified reference GraphCNF from "
                                          "baseline/_ref (bisection inverse), %.1f s of CPU work" % (reps, nsb, total)}
    return out4, out5


def run_all(args, rank, world, dev, dist, clock_sampler=None):
    if not hasattr(args, "no_train"):
        args.no_train = False
    c = Ctx(args, rank, world, dev, dist)
    records = []
    for fn in (run_graph_coloring, run_molecules):
        try:
            if clock_sampler is not None:
                clock_sampler.start()
            r = fn(c)
            r = r if isinstance(r, list) else [r]
            if clock_sampler is not None:
                clk = clock_sampler.stop()
                for rec in r:
                    rec["clocks"] = clk
            records += r
        except Exception as exc:      # noqa: BLE001 - a failing extra config must not take the headline line down with it
            import traceback
            records.append({"name": fn.__name__, "error": "%s: %s" % (type(exc).__name__, exc),
                            "traceback": traceback.format_exc()[-1500:]})
            if c.distributed:
                raise                  # ranks would desynchronise: fail loudly
        torch.cuda.empty_cache()
    return records
