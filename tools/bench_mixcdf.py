"""Kernel-level timing of cnf_mixcdf_fwd / cnf_mixcdf_inv at the LM shape (B=4096, S=256, C=16, K=8).
python tools/bench_mixcdf.py [--inv] [--B 4096] [--K 8] [--reps 20]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096); ap.add_argument("--S", type=int, default=256)
ap.add_argument("--C", type=int, default=16); ap.add_argument("--K", type=int, default=8)
ap.add_argument("--reps", type=int, default=20); ap.add_argument("--inv", action="store_true")
ap.add_argument("--std", type=float, default=0.5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, C, K = a.B, a.S, a.C, a.K
g = torch.Generator(device=dev).manual_seed(0)
nbuf = 3
nn = [torch.randn(B, S, C * (2 + 3 * K), device=dev, generator=g) * a.std for _ in range(nbuf)]
z = torch.randn(B, S, C, device=dev, generator=g) * (1.0 / 1.81 * 1.8)
sf = torch.randn(C, device=dev, generator=g) * 0.3
msf = torch.randn(C, K, device=dev, generator=g) * 0.3
mask_c = [1.0] * (C // 2) + [0.0] * (C - C // 2)
out = torch.empty_like(z)
def run(i):
    return ops.mixcdf(z, nn[i % nbuf], K, mask_c=mask_c, scaling_factor=sf, mixture_scaling_factor=msf, reverse=a.inv, out=out)
for i in range(3): run(i)
torch.cuda.synchronize()
evs = []
for i in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(i); e1.record(); evs.append((e0, e1))
torch.cuda.synchronize()
ms = sorted(x.elapsed_time(y) for x, y in evs)
Ct = C - C // 2
alg = B * S * (8 * C + 4 * Ct * (2 + 3 * K))
med = ms[len(ms) // 2]
ops.check_status(dev)
print(json.dumps({"kernel": "mixcdf_inv" if a.inv else "mixcdf_fwd", "generic": bool(os.environ.get("CNF_B200_MIXCDF_GENERIC")),
                  "B": B, "S": S, "C": C, "K": K, "ms_median": med, "ms_min": ms[0], "ms_max": ms[-1],
                  "GBps": alg / med / 1e6, "samples_per_s": B / med * 1e3}))
