"""Time cnf_categ_encode at the LM shape (B 4096, S 256, V 51, d 16), with and without the fused first block.
    python tools/bench_encode.py [--reps 20]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workload as W
from categoricalnf_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
prm = W.lm_params(seed=0)
path = W.LMDevicePath(prm, dev)
B, S = W.LM["B"], W.LM["S"]
tokens = W.lm_tokens(B, S, prm.V).to(dev)
nxt = (path.blocks[0]["bias"], path.blocks[0]["scales"], path.blocks[0]["w"])
for name, kw in (("plain", {}), ("fused_first_block", dict(fuse_next=nxt))):
    ldj = torch.zeros(B, device=dev)
    fn = lambda: ops.categ_encode(tokens, path.table, path.prior, ldj, seed=1, **kw)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print('{"kernel": "categ_encode_%s", "ms": %.4f, "tokens_per_s": %.0f}' % (name, ms, B * S / ms * 1e3))
