"""Training step of the LM workload (BASELINE config 2 shape: B 4096, S 256, d 16, K 8, 8 blocks, stand-in Linear coupling
nets) through the drop-in modules: forward in training mode, loss = mean bits/dim, backward through the hand-written backward
kernels.  Prints ms per step (forward / forward+backward) and samples/s; --profile adds a kernel table;
--cpu times the same step through autograd over the CPU oracle on a 128-sample batch.
    python tools/bench_train.py [--batch 4096] [--reps 5] [--profile] [--cpu]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workload as W
from categoricalnf_b200 import functional as CF

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=W.LM["B"])
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--cpu", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, V = a.batch, W.LM["S"], W.LM["V"]
prm = W.lm_params(seed=0)
W.LMDevicePath(prm, dev).data_init(seed=0)
model, _ = W.build_lm_model(prm, dev)
model.train()
tokens = W.lm_tokens(B, S, V, seed=0).to(dev)
params = [p for p in model.parameters() if p.requires_grad]


def fwd():
    z, ldj = model(tokens)
    lp = CF.logistic_logprob(z).sum(dim=[1, 2])
    return ((-ldj - lp) / S).mean()


def step():
    for p in params:
        p.grad = None
    loss = fwd()
    loss.backward()
    return loss


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


t_f = timed(fwd, a.reps)
t_s = timed(step, a.reps)
loss = step()
gnorm = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in params if p.grad is not None)).item()
out = {"workload": "LM training step B=%d S=%d d=16 K=8 x 8 blocks" % (B, S), "fwd_train_ms": t_f, "fwd_bwd_ms": t_s,
       "samples_per_s": B / t_s * 1e3, "loss_bpd": loss.item() * 1.4426950408889634, "grad_norm": gnorm,
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
if a.cpu:
    from oracle import cnf_oracle as O
    nb = 128
    tok, u = W.lm_tokens(nb, S, V, seed=1), W.lm_noise(nb, S, prm.D, seed=1)
    leaf = lambda t: t.clone().requires_grad_(True)
    table = torch.nn.functional.linear(leaf(prm.embed_w), leaf(prm.pred_w), leaf(prm.pred_b))
    blocks = []
    for b in prm.blocks:
        lv = {k: leaf(b[k]) for k in ("bias", "scales", "l", "u", "log_s", "net_w", "net_b", "sf", "msf")}
        w, sldj = O.invconv_weight(b["p"], lv["l"], lv["log_s"], lv["u"], b["sign_s"])
        blocks.append(dict(bias=lv["bias"].view(1, 1, -1), scales=lv["scales"].view(1, 1, -1), weight=w, sldj=sldj, mask=b["mask"],
                           K=prm.K, sf=lv["sf"], msf=lv["msf"],
                           nn_fn=(lambda zin, lv=lv: torch.nn.functional.linear(zin, lv["net_w"], lv["net_b"]))))
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    z, ldj, lp = O.lm_flow_forward(tok, u, dict(table=table, prior=prm.prior), blocks)
    ((-ldj - lp) / S).mean().backward()
    dt = time.perf_counter() - t0
    out.update(cpu_oracle_samples_per_s=nb / dt, cpu_cores=os.cpu_count(), cpu_sample="one fwd+bwd at B=128")
print(json.dumps(out))
if a.profile:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=64))
