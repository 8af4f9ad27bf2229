#!/usr/bin/env python
"""bench.py - flow forward + log-det-Jacobian throughput on BASELINE.json's LM configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (configs[1]): tokens [4096, 256] over 51 classes -> mixture-of-logistics encoding (d=16)
-> 8 x [ActNorm, 1x1 InvertibleConv, MixtureCDFCoupling K=8] -> logistic prior log-prob.  One step =
one pass of that hot path over one batch; every rank runs its own batch of 4096 (weak scaling) and
the ranks all-reduce (sum log-likelihood, sample count) once per step.

`value`  kernel-level: coupling-network outputs given as HBM-resident inputs (8 x 1.74 GB, far
         larger than L2), tokens resident, device-timed with CUDA events.
`e2e`    module-level: the drop-in FlowModel (stand-in Linear coupling nets evaluated on the
         device) driven from pinned HOST tokens, H2D copy and D2H read of the per-sample
         log-likelihood inside the timed region.
`roofline` the dominant kernel (mixture coupling forward): algorithmic bytes per launch / mean
         launch duration from CUDA events recorded inside the timed region.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference modules (baseline/_ref, vendored by
         tools/vendor_reference.sh) on the host cores, on a bounded sub-batch of the same workload;
         the oracle port only when baseline/_ref did not travel with the repo.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import workload as W  # noqa: E402

METRIC = "flow fwd+ldj samples/sec"
UNIT = "samples/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.index = None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference modules from baseline/_ref (tools/vendor_reference.sh) when they travelled with the
# repo (kind "reference"), else the oracle port of the reference's eager fp64 path (kind "port")
# --------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, prm):
        self.prm = prm
        self.kind = "reference" if W.reference_available() else "port"
        if self.kind == "reference":
            self.model, self.prior = W.build_lm_reference_model(prm)
        self.what = ("unmodified reference modules (baseline/_ref: layers/flows, layers/categorical_encoding; eager fp32/fp64)"
                     if self.kind == "reference" else "eager fp64 oracle port (baseline/_ref absent)")

    def step(self, B, seed):
        prm = self.prm
        tokens = W.lm_tokens(B, prm.S, prm.V, seed=seed)
        if self.kind == "reference":
            torch.manual_seed(seed)
            t0 = time.perf_counter()
            with torch.no_grad():
                z, ldj = self.model(tokens, reverse=False)
                logp = self.prior.log_prob(z).sum(dim=[1, 2])
        else:
            u = W.lm_noise(B, prm.S, prm.D, seed=seed)
            t0 = time.perf_counter()
            z, ldj, logp = W.lm_oracle_forward(prm, tokens, u)
        return time.perf_counter() - t0, W.bits_per_dim(ldj, logp, prm.S)

    def pick_batch(self, steps, budget_s):
        """Sub-batch so that `steps` CPU steps fit in about `budget_s` seconds (calibrated on B=8)."""
        self.step(4, 99)
        dt, _ = self.step(8, 98)
        per_sample = dt / 8
        B = 8
        while B < 512 and 2 * B * per_sample * steps <= budget_s:
            B *= 2
        return B

    def cross_check(self, B=4):
        """reference arm only: the oracle port on the noise the reference drew - how far the test oracle is from the real thing."""
        if self.kind != "reference":
            return None
        tokens = W.lm_tokens(B, self.prm.S, self.prm.V, seed=55)
        z, ldj, lp, u = W.lm_reference_forward(self.model, self.prior, tokens, 55)
        z2, ldj2, lp2 = W.lm_oracle_forward(self.prm, tokens, u)
        return {"batch": B, "max_abs_z": float((z - z2).abs().max()), "max_rel_ldj": float(((ldj - ldj2).abs() / ldj2.abs()).max())}


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    prm = W.data_init_oracle(W.lm_params(seed=0), seed=0)
    arm = CpuArm(prm)
    B = arm.pick_batch(args.steps + args.warmup, budget_s=150.0)
    for i in range(args.warmup):
        arm.step(B, 100 + i)
    times, bpd = [], None
    for i in range(args.steps):
        dt, bpd = arm.step(B, i)
        times.append(dt)
    total = sum(times)
    value = B * args.steps / total
    sample = "B=%d of the %d-sample batch per step (S=%d, d=%d, K=%d, %d blocks), %s" % (
        B, W.LM["B"], prm.S, prm.D, prm.K, len(prm.blocks), arm.what)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": arm.kind, "sample": sample,
                         "sub_batch": B},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "bits_per_dim": bpd, "oracle_vs_reference": arm.cross_check(),
    }
    print(json.dumps(line), flush=True)


def same_on_all_ranks(n, dev, dist):
    """A step count derived from a LOCAL timing must be agreed on before it drives a loop that issues collectives: every rank
    takes the maximum (ranks with different counts would wait for each other's all-reduces forever)."""
    if dist is None:
        return int(n)
    t = torch.tensor([int(n)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def workload_config(n_gpus, extra=None):
    cfg = {"workload": "language_modeling: synthetic char-level tokens, seq 256, d=16, V=51, 8 x [ActNorm, InvConv1x1, "
                       "MixtureCDFCoupling K=8], batch 4096 per GPU",
           "batch_per_gpu": W.LM["B"], "global_batch": W.LM["B"] * n_gpus, "seq_len": W.LM["S"], "d": W.LM["D"],
           "mixtures": W.LM["K"], "vocab": W.LM["V"], "blocks": W.LM["blocks"], "parallelism": "batch-sharded x%d" % n_gpus,
           "l2": "inputs larger than L2 (1.74 GB of coupling parameters per layer vs 126 MB L2); no flush needed",
           "value_leg": "coupling-net outputs given, resident in HBM", "e2e_leg": "drop-in FlowModel (see e2e.mode), stand-in Linear "
           "coupling nets (final projection fused with the mixture transform on tcgen05, 3xTF32), pinned host tokens -> H2D "
           "(double-buffered on a copy stream), per-sample log-likelihood + kernel status word -> D2H (host reads step i-1 while step i runs)"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device - the product path has no CPU fallback "
                           "(use --impl reference for the CPU arm)")
    from categoricalnf_b200 import ops
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    B, S, D, K, V = (W.LM[k] for k in ("B", "S", "D", "K", "V"))
    prm = W.lm_params(seed=0)
    path = W.LMDevicePath(prm, dev).data_init(seed=0)
    tokens = W.lm_tokens(B, S, V, seed=rank).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    nn_outs = [torch.randn(B, S, D * (2 + 3 * K), device=dev, generator=gen) * 0.5 for _ in prm.blocks]
    from categoricalnf_b200.sharding import LogLikAllReducer
    # (sum log-likelihood, count) of every step: produced by the epilogue of the step's last kernel (the prior log-prob, which
    # also adds ldj), all-reduced over the ranks on a communication stream - nothing is queued on the compute stream for it
    reducer = LogLikAllReducer(dev, slots=4)

    def barrier():
        reducer.finish()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, time_mix=False):
        z, ldj, ll = path.forward(tokens, nn_outs=nn_outs, seed=rank, offset=i * B * S * D, time_mix=time_mix,
                                  total=reducer.slot())
        reducer.reduce()
        return ldj, ll

    for i in range(args.warmup):
        step(i)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    path.mix_events = []
    barrier()
    ev0.record()
    for i in range(args.steps):
        ldj, ll = step(args.warmup + i, time_mix=True)
    reducer.finish()          # the timed region ends when the last step's all-reduce has landed
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ops.launch_count() - launches0
    mix_ms = [a.elapsed_time(b) for a, b in path.mix_events]
    bpd_gpu = W.bits_per_dim(ll, torch.zeros_like(ll), S)
    pair = reducer.result(reducer.step - 1)          # global (sum log-likelihood, count) of the last step
    bpd_global = float(-pair[0] / (pair[1] * S) * 1.4426950408889634)
    ops.check_status(dev, "bench value leg")

    # ---- e2e: module API from pinned host tokens -------------------------------------------------
    # (nn_outs - 14 GB of 180 - stay resident: the sustained value leg runs after both timed regions)
    model, prior = W.build_lm_model(prm, dev)
    parity = parity_check(prm, model, dev) if rank == 0 else None
    host_tokens = [W.lm_tokens(B, S, V, seed=10 * rank + j).pin_memory() for j in range(2)]
    host_ll = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_tokens = [torch.empty(B, S, dtype=torch.int64, device=dev) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]      # tokens of slot j have landed
    consumed = [torch.cuda.Event() for _ in range(2)]   # the step that read slot j has been issued and finished with it
    result = [torch.cuda.Event() for _ in range(2)]     # log-likelihoods of slot j are in pinned host memory

    # Double-buffered input pipeline: the H2D copy of step i+1 runs on a copy stream while step i computes, and the host
    # waits for the D2H result of step i-1 while step i is in flight.  Every step's H2D and D2H happen inside the timed
    # region (the first copy is not overlapped, the last result is awaited before the closing event).
    def prefetch(i):
        j = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[j])
            dev_tokens[j].copy_(host_tokens[j], non_blocking=True)
            ready[j].record(copy_stream)

    graphed = None
    if not args.eager_e2e:
        from categoricalnf_b200.layers.flows import GraphedFlowForward
        pair_static = torch.zeros(2, dtype=torch.float64, device=dev)      # written by the graph's last kernel every replay
        graphed = GraphedFlowForward(model, log_likelihood=lambda z, ldj, pad: ops.logistic_logprob(z, pad=pad, add=ldj,
                                                                                                     total=pair_static)[0])
    if graphed is not None:
        try:      # capture now; a box where the capture fails still gets an e2e number (launch by launch)
            dev_tokens[0].copy_(host_tokens[0])
            graphed(dev_tokens[0])
            torch.cuda.synchronize()
        except Exception as exc:      # noqa: BLE001
            print("bench: CUDA-graph replay unavailable (%s: %s) - the e2e leg launches kernel by kernel" % (type(exc).__name__, exc),
                  file=sys.stderr)
            graphed = None
    e2e_mode = "replayed from a CUDA graph (GraphedFlowForward)" if graphed is not None else "launched kernel by kernel"
    status_dev = ops.status_word(dev)
    host_status = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]

    def consume(j):
        result[j].synchronize()
        if int(host_status[j][0]) != 0:
            ops.check_status(dev, "bench e2e leg")      # raises like the reference's per-layer asserts would have

    def e2e_step(i, last):
        j = i % 2
        cur = torch.cuda.current_stream()
        if not last:
            prefetch(i + 1)
        with torch.no_grad():
            cur.wait_event(ready[j])
            # numerical-health word of the kernels (NaN / CDF-range flags): instead of one blocking read per forward
            # (check_nan=True) it travels to the host with the step's result and is examined when that result is consumed
            if graphed is not None:
                # whole forward + prior log-likelihood replayed from one CUDA graph (tokens copied into its static buffer)
                slot = reducer.slot()
                z, ldj, ll = graphed(dev_tokens[j])
                consumed[j].record(cur)
                if distributed:
                    slot.copy_(pair_static, non_blocking=True)
            else:
                z, ldj = model(dev_tokens[j], check_nan=False)
                consumed[j].record(cur)
                ll, _ = ops.logistic_logprob(z, add=ldj, total=reducer.slot())
            reducer.reduce()                      # 16-byte all-reduce on the communication stream (no-op at N = 1)
            host_ll[j].copy_(ll, non_blocking=True)
            host_status[j].copy_(status_dev, non_blocking=True)
            result[j].record(cur)
        if i > 0:
            consume((i - 1) % 2)                  # the host consumes the previous step's log-likelihoods
        if last:
            consume(j)

    def e2e_run(n):
        for j in range(2):
            consumed[j].record(torch.cuda.current_stream())
        prefetch(0)
        for i in range(n):
            e2e_step(i, i == n - 1)

    e2e_run(args.warmup)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(args.steps)
    reducer.finish()
    e1.record()
    barrier()
    e2e_ms_total = e0.elapsed_time(e1)
    ops.check_status(dev, "bench e2e leg")
    clk = clocks.stop() if rank == 0 else None

    # ---- both legs sustained for >= 1 s each (the timed regions above are K steps = tens of milliseconds).  They run AFTER
    # the timed regions and under their own clock sampler: a second of these kernels back to back reaches the board's power
    # cap, which must neither throttle the timed e2e region nor be mixed into its clock record.
    sus_clocks = ClockSampler(local_rank)
    if rank == 0:
        sus_clocks.start()
    sus_n = same_on_all_ranks(max(args.steps, int(1100.0 / (ms_total / args.steps)) + 1), dev, dist if distributed else None)
    barrier()
    ev0.record()
    for i in range(sus_n):
        step(args.warmup + args.steps + i)
    reducer.finish()
    ev1.record()
    barrier()
    sus_ms = ev0.elapsed_time(ev1)
    ops.check_status(dev, "bench value leg (sustained)")
    del nn_outs
    torch.cuda.empty_cache()
    e2e_sus_n = same_on_all_ranks(max(args.steps, int(1100.0 / (e2e_ms_total / args.steps)) + 1), dev, dist if distributed else None)
    e0.record()
    e2e_run(e2e_sus_n)
    reducer.finish()
    e1.record()
    barrier()
    e2e_sus_ms = e0.elapsed_time(e1)
    ops.check_status(dev, "bench e2e leg (sustained)")
    sus_clk = sus_clocks.stop() if rank == 0 else None

    # ---- max over ranks ---------------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_ms_total, sum(mix_ms) / max(1, len(mix_ms)), sus_ms, e2e_sus_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total, mix_ms_mean, sus_ms, e2e_sus_ms = (float(x) for x in t.tolist())

    # ---- training step of the same flow (the backward kernels), weak scaling ------------------------------
    lm_train = None
    if not args.no_train:
        try:
            lm_train = lm_training_record(args, model, prm, dev, rank, world, dist if distributed else None, tokens)
        except Exception as exc:      # noqa: BLE001 - an extra record must not take the headline line down with it
            import traceback
            lm_train = {"name": "language_modeling_train", "error": "%s: %s" % (type(exc).__name__, exc),
                        "traceback": traceback.format_exc()[-1500:]}
            if distributed:
                raise

    # ---- BASELINE configs 3, 4, 5 (graph colouring, GraphCNF log-likelihood, GraphCNF sampling) ---------------------------
    graph_records = None
    if not args.no_graphs:
        del model, prior, graphed
        torch.cuda.empty_cache()
        import bench_graphs
        graph_records = bench_graphs.run_all(args, rank, world, dev, dist, ClockSampler(local_rank) if rank == 0 else None)

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = ms_total / args.steps
        value = B * world * args.steps / (ms_total * 1e-3)
        e2e_value = B * world * args.steps / (e2e_ms_total * 1e-3)
        Ct = D - D // 2
        alg_bytes = B * S * (4 * D + 4 * D + 4 * Ct * (2 + 3 * K))     # 960 B / position
        achieved = alg_bytes / (mix_ms_mean * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * S * 8, "d2h_bytes_per_step": B * 4 + 4,
                    "ms_per_step": e2e_ms_total / args.steps, "mode": e2e_mode,
                    "sustained": {"steps": e2e_sus_n, "seconds": e2e_sus_ms * 1e-3,
                                  "value": B * world * e2e_sus_n / (e2e_sus_ms * 1e-3), "ms_per_step": e2e_sus_ms / e2e_sus_n}},
            "sustained": {"steps": sus_n, "seconds": sus_ms * 1e-3, "value": B * world * sus_n / (sus_ms * 1e-3),
                          "ms_per_step": sus_ms / sus_n, "clocks": sus_clk,
                          "note": "the value leg's step (and, under e2e.sustained, the e2e step) repeated for >= 1 s after both "
                                  "timed regions, under its own clock sampler"},
            "gpu_launches": launches,
            "roofline": {"kernel": "mixcdf_pipe_kernel<8,8,fwd> (cnf_mixcdf_fwd)", "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "mean_launch_ms": mix_ms_mean, "launches_timed": len(mix_ms),
                         "share_of_step": mix_ms_mean * len(prm.blocks) / ms_step},
            "clocks": clk, "bits_per_dim": bpd_gpu, "bits_per_dim_all_ranks": bpd_global, "parity": parity,
            "collective": "one all-reduce of (sum log-likelihood, count) = 16 bytes per step; the pair comes from the epilogue of "
                          "the step's last kernel (cnf_logistic_logprob add/total) and is reduced on a communication stream",
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(prm)
        records = ([lm_train] if lm_train is not None else []) + (graph_records or [])
        if records:
            line["configs"] = records
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def lm_training_record(args, model, prm, dev, rank, world, dist, tokens):
    """Training step of the headline flow at its shape (4096 samples per GPU, weak scaling): forward in training mode through
    the drop-in modules, loss = mean bits/dim, backward through the hand-written backward kernels (cnf_*_bwd), gradients
    all-reduced in flat buckets on a communication stream while backward runs (sharding.GradientReducer), fused Adam.
    The reference runs this step by autograd over its eager modules (general/train.py:148-160)."""
    from categoricalnf_b200 import functional as CF
    from categoricalnf_b200 import ops
    from categoricalnf_b200.sharding import GradientReducer
    B, S = tokens.shape
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    red = GradientReducer(params, bucket_bytes=32 << 20, profile=True)
    opt = torch.optim.Adam(params, lr=1e-7, fused=True)
    state = {}

    def train_step():
        red.zero_grad()
        z, ldj = model(tokens)
        lp = CF.logistic_logprob(z).sum(dim=[1, 2])
        loss = ((-ldj - lp) / S).mean()
        loss.backward()
        red.finish()
        opt.step()
        state["loss"] = loss.detach()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps, warmup = max(2, min(args.steps, 5)), 3
    try:
        train_step()
        n0 = ops.launch_count()
        train_step()
        launches = ops.launch_count() - n0
        for _ in range(warmup - 2):
            train_step()
        barrier()
        red.comm_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            train_step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        nbytes, comm_ms, bus = red.comm_stats()
        ops.check_status(dev, "bench language_modeling training")
        return {"name": "language_modeling_train", "baseline_config": "configs[1], training step", "unit": "samples/s",
                "metric": "LM flow training step (fwd + bwd + gradient all-reduce + Adam) samples/sec",
                "value": B * world / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": warmup, "scaling": "weak",
                "batch_per_gpu": B, "global_batch": B * world, "parameters": sum(p.numel() for p in params),
                "gpu_launches_per_step": launches, "loss_bits_per_dim": float(state["loss"]) * 1.4426950408889634,
                "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30, "dtype": "f32 (projections 3xTF32)",
                "mode": "eager; stand-in Linear coupling nets with the compact [B,S,Ct(2+3K)] projection (only the transformed "
                        "channels' weight rows, coupling mask folded into the weight block); backward: the projection's grad_x / "
                        "grad_W / grad_b are formed inside the mixture transform's backward kernel from the gradient tile in "
                        "shared memory (dL/dnn_out is never stored), ActNorm / 1x1 conv / encode backward kernels; gradients "
                        "reduced in %d flat bucket(s) on a communication stream, fused Adam" % len(red.buckets),
                "collective": {"kind": "NCCL all-reduce (sum) of the flat gradient buckets, overlapped with backward",
                               "bytes_per_step": nbytes / steps if world > 1 else 0,
                               "comm_stream_ms_per_step": comm_ms / steps if world > 1 else 0.0, "bus_GBps": bus}}
    finally:
        red.close()
        model.eval()
        for p in params:
            p.grad = None


def parity_check(prm, model, dev, B=32):
    """Same tokens and noise through the drop-in modules (GPU) and the CPU oracle: bits/dim of both and
    the worst tolerance-normalised deviation of z and ldj (<= 1 means inside |a-b| <= 1e-4|b| + atol)."""
    from categoricalnf_b200 import ops
    tokens, u = W.lm_tokens(B, prm.S, prm.V, seed=77), W.lm_noise(B, prm.S, prm.D, seed=77)
    z_ref, ldj_ref, lp_ref = W.lm_oracle_forward(prm, tokens, u)
    with torch.no_grad():
        z, ldj = model(tokens.to(dev), u_noise=u.to(dev))
        lp, _ = ops.logistic_logprob(z)
    z, ldj, lp = z.cpu().double(), ldj.cpu().double(), lp.cpu().double()
    dev_z = ((z - z_ref).abs() / (1e-4 * z_ref.abs() + 1e-5)).max().item()
    dev_ldj = ((ldj - ldj_ref).abs() / (1e-4 * ldj_ref.abs() + 2e-4)).max().item()
    return {"batch": B, "bits_per_dim_gpu": W.bits_per_dim(ldj, lp, prm.S),
            "bits_per_dim_oracle": W.bits_per_dim(ldj_ref, lp_ref, prm.S), "z_dev_over_tol": dev_z,
            "ldj_dev_over_tol": dev_ldj}


def ncu_traffic():
    """dram bytes read+written per mixture-coupling launch from the committed ncu --set full capture
    (profiles/*.json written by tools/ncu_summary.py), or None."""
    path = os.path.join(ROOT, "profiles", "mixcdf_fwd_traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["dram_bytes_per_launch"])
        except (ValueError, KeyError):
            return None
    return None


def cpu_baseline(prm):
    """The reference (baseline/_ref) - or, when it did not travel, the oracle port - on the host cores, on a bounded
    sub-batch (about 10-30 s of CPU work)."""
    torch.set_num_threads(os.cpu_count() or 1)
    arm = CpuArm(prm)
    B = arm.pick_batch(1, budget_s=20.0)
    arm.step(B, 1)                           # warm-up at the measured size (allocator, thread pool)
    times, bpd, t_start = [], None, time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < 12.0 and len(times) < 8):
        dt, bpd = arm.step(B, len(times))
        times.append(dt)
    total = sum(times)
    return {"value": B * len(times) / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": arm.kind,
            "sample": "%d steps at B=%d of the %d-sample batch (S=%d, d=%d, K=%d, %d blocks, stand-in Linear nets), "
                      "%s, %.1f s of CPU work, best step %.0f samples/s"
                      % (len(times), B, W.LM["B"], prm.S, prm.D, prm.K, len(prm.blocks), arm.what, total, B / min(times)),
            "bits_per_dim": bpd}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / reference-parity legs")
    ap.add_argument("--no-graphs", action="store_true", help="skip the records of BASELINE configs 3, 4, 5 (key `configs`)")
    ap.add_argument("--no-train", action="store_true", help="skip the GraphCNF training-step record")
    ap.add_argument("--graph-steps", type=int, default=5, help="timed steps per graph config (capped by --steps)")
    ap.add_argument("--eager-e2e", action="store_true", help="e2e leg launches kernel by kernel instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
