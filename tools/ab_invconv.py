"""A/B of the invertible 1x1 convolution at the LM shape (z [4096, 256, 16] @ W [16, 16], reference
layers/flows/permutation_layers.py:106-136):

  A  cnf_invconv_apply   - invconv_rows_kernel<16>: one thread per position on the CUDA cores, row in registers, W broadcast
                           from shared memory (the adopted kernel); also with the ActNorm prologue + masked second output
                           that the e2e step uses
  B  cnf_linear_fwd      - the tcgen05 projection kernel on the same product (M = positions, N = K = 16): TMA loads into
                           128-byte-swizzled shared memory, tcgen05.mma kind::tf32 (3xTF32 for fp32 accuracy, and one-pass TF32
                           as the speed bound of this design), accumulator in tensor memory, TMA store - the north_star's
                           "small-K per-position GEMM on tensor cores fed by TMA"
  C  torch copy          - z.clone(): the measured read+write bandwidth bound of this pass (134 MB)

Each variant: median of --reps launches timed one by one with CUDA events, inputs rotated over 3 buffers (3 x 67 MB + outputs
>> L2 share is small but non-zero; the ordering of A / B / C is unaffected).  Prints one JSON line.
    python tools/ab_invconv.py [--reps 30]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--S", type=int, default=256)
ap.add_argument("--C", type=int, default=16)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--only", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, C = a.B, a.S, a.C
g = torch.Generator(device=dev).manual_seed(0)
zs = [torch.randn(B, S, C, device=dev, generator=g) for _ in range(3)]
w = torch.linalg.qr(torch.randn(C, C, device=dev, generator=g))[0].contiguous()
wt = w.t().contiguous()                      # nn.Linear weight: y = x @ weight.T
sldj = torch.zeros(1, device=dev)
bias, scales = torch.randn(C, device=dev, generator=g) * 0.1, torch.randn(C, device=dev, generator=g) * 0.1
omask = torch.tensor([1.0] * (C // 2) + [0.0] * (C - C // 2), device=dev)
ldj = torch.zeros(B, device=dev)


def timed(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    evs = []
    for i in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(x.elapsed_time(y) for x, y in evs)
    return ms[len(ms) // 2], ms[0]


variants = {
    "A_rows_plain": (lambda i: ops.invconv_apply(zs[i % 3], w, sldj, None), 2),
    "A_rows_actnorm_masked": (lambda i: ops.invconv_apply(zs[i % 3], w, sldj, None, pre_actnorm=(bias, scales), out_mask=omask), 3),
    "B_tcgen05_3xtf32": (lambda i: ops.linear(zs[i % 3].view(-1, C), wt, None, precision="3xtf32"), 2),
    "B_tcgen05_tf32": (lambda i: ops.linear(zs[i % 3].view(-1, C), wt, None, precision="tf32"), 2),
    "C_copy": (lambda i: zs[i % 3].clone(), 2),
}
out = {"shape": [B, S, C], "reps": a.reps, "bytes_per_pass": B * S * C * 4}
ref = zs[0].double().view(-1, C) @ w.double()
for name, (fn, passes) in variants.items():
    if a.only and a.only not in name:
        continue
    med, best = timed(fn)
    nbytes = passes * B * S * C * 4
    out[name] = {"ms_median": med, "ms_min": best, "GBps": nbytes / med / 1e6, "bytes": nbytes}
# numerics of the two candidates against float64
ya = ops.invconv_apply(zs[0], w, sldj, None)[0].double().view(-1, C)
yb = ops.linear(zs[0].view(-1, C), wt, None, precision="3xtf32").double()
yc = ops.linear(zs[0].view(-1, C), wt, None, precision="tf32").double()
out["max_abs_err_vs_f64"] = {"A_rows": float((ya - ref).abs().max()), "B_3xtf32": float((yb - ref).abs().max()),
                             "B_tf32": float((yc - ref).abs().max())}
print(json.dumps(out))
