"""Time cnf_linear_mixcdf_fwd at the LM shape (B 4096, S 256, C 16, K 8) for a given in_features H.
    python tools/bench_fused.py [--H 16] [--reps 10] [--prec 3xtf32] [--inv]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=16)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--prec", default="3xtf32")
ap.add_argument("--inv", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, C, K = 4096, 256, 16, 8
PN = 2 + 3 * K
z = torch.randn(B, S, C, device=dev)
feats = torch.randn(B, S, a.H, device=dev)
w = torch.nn.Parameter(torch.randn(C * PN, a.H, device=dev) * (0.5 / a.H ** 0.5))      # parameters: folded-bias path
b = torch.nn.Parameter(torch.randn(C * PN, device=dev) * 0.1)
mask_c = [1.0] * 8 + [0.0] * 8
fn = lambda: ops.linear_mixcdf(z, feats, w, b, K, mask_c=mask_c, precision=a.prec, reverse=a.inv)
torch.set_grad_enabled(False)
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
print('{"kernel": "linear_mixcdf_%s", "H": %d, "precision": "%s", "ms": %.4f, "samples_per_s": %.0f}'
      % ("inv" if a.inv else "fwd", a.H, a.prec, ms, B / ms * 1e3))
