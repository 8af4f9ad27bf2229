"""Time the mixture-coupling backward kernel (cnf_mixcdf_bwd) and a full training-direction pass of one
coupling (forward + backward) at the LM shape.   python tools/bench_bwd.py [--B 4096] [--reps 5]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops, ops_bwd
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4096)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, S, C, K = a.B, 256, 16, 8
PN = 2 + 3 * K
z = torch.randn(B, S, C, device=dev)
nn_out = torch.randn(B, S, C * PN, device=dev) * 0.5
sf, msf = torch.randn(C, device=dev) * 0.3, torch.randn(C, K, device=dev) * 0.3
gz, gl = torch.randn(B, S, C, device=dev), torch.randn(B, device=dev)
cfg = dict(K=K, mask_c=[1.0] * 8 + [0.0] * 8, mask_s=None, reverse=False, reg_max=-1.0, reg_factor=1.0, training=True, prebounded=False)


def timeit(fn, reps):
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


t_b = timeit(lambda: ops_bwd.mixcdf_backward(cfg, z, nn_out, sf, msf, None, None, gz, gl, None), a.reps)
t_f = timeit(lambda: ops.mixcdf(z, nn_out, K, mask_c=cfg["mask_c"], scaling_factor=sf, mixture_scaling_factor=msf), a.reps)
P = B * S
byts = P * (4 * C * 3 + 4 * 8 * PN + 4 * C * PN)   # z, gz_out in, gz out; params of transformed channels in; full grad rows out
print('{"kernel": "mixcdf_bwd", "B": %d, "ms": %.4f, "GBps": %.0f, "fwd_ms": %.4f, "algorithmic_bytes": %d}'
      % (B, t_b, byts / t_b / 1e6, t_f, byts))
