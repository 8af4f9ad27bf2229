"""GPU check of cnf_linear_fwd (tcgen05 Linear): error against a float64 product for a sweep of shapes,
both precisions, both store paths (N % 4 != 0 takes the direct-store epilogue), then timings.
    python tools/linear_check.py [tf32|3xtf32] [--time]
Prints one line per case; exit code 1 if any case is out of tolerance."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops

prec = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "3xtf32"
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
# relative-to-scale tolerance: |err| <= tol * sqrt(K) * rms(x) * rms(w)
TOL = {"tf32": 2e-3, "3xtf32": 4e-6}[prec]
cases = [
    (128, 32, 32, True, None), (128, 32, 8, True, None), (100, 32, 64, True, None), (1000, 16, 416, True, None),
    (4096, 384, 208, True, None), (300, 100, 50, True, None), (65536, 64, 512, False, None), (5000, 64, 256, True, "gelu"),
    (777, 36, 418, True, None), (128 * 148 * 2 + 5, 128, 96, True, "gelu"), (33, 1024, 1300, True, None),
]
bad = 0
for (M, K, N, use_bias, act) in cases:
    g = torch.Generator(device="cpu").manual_seed(M + 7 * K + 13 * N)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev) if use_bias else None
    y = ops.linear(x, w, b, precision=prec, activation=act)
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    if b is not None:
        ref = ref + b.double()
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    err = (y.double() - ref).abs().max().item()
    scale = 1.0   # rms(x)=1, rms(w)=1/sqrt(K) -> product terms sum to O(1)
    ok = err <= TOL * scale * 4 and bool(torch.isfinite(y).all())
    bad += (not ok)
    print("%-7s M=%-7d K=%-5d N=%-5d bias=%d act=%-5s max|err|=%.3e  %s" % (prec, M, K, N, use_bias, act, err, "ok" if ok else "FAIL"),
          flush=True)

if "--time" in sys.argv:
    def timeit(fn, reps=20):
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
    for (M, K, N) in [(4096 * 256, 16, 416), (512 * 703, 192, 384), (1024 * 20 * 8, 384, 1536), (65536, 1024, 1024)]:
        x = torch.randn(M, K, device=dev)
        w = torch.randn(N, K, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        t_tc = timeit(lambda: ops.linear(x, w, b, precision=prec))
        if prec == "3xtf32":
            wp = w.clone()
            wp._cnf_cache_lo = True          # weight split precomputed and cached (what nn.Parameters get)
            t_pre = timeit(lambda: ops.linear(x, wp, b, precision=prec))
            t_pre128 = timeit(lambda: ops.linear(x, wp, b, precision=prec, block_n=128))
            t_128 = timeit(lambda: ops.linear(x, w, b, precision=prec, block_n=128))
            print("     3xtf32 variants M=%d K=%d N=%d: in-kernel split %.3f ms | BN=128 %.3f | pre-split W %.3f | pre-split W, BN=128 %.3f"
                  % (M, K, N, t_tc, t_128, t_pre, t_pre128), flush=True)
        torch.backends.cuda.matmul.allow_tf32 = False
        t_fp32 = timeit(lambda: torch.nn.functional.linear(x, w, b))
        torch.backends.cuda.matmul.allow_tf32 = True
        t_tf32 = timeit(lambda: torch.nn.functional.linear(x, w, b))
        torch.backends.cuda.matmul.allow_tf32 = False
        flops = 2.0 * M * N * K
        byts = 4.0 * (M * K + N * K + M * N)
        print("time %-7s M=%d K=%d N=%d: tcgen05 %.3f ms (%.1f TFLOP/s, %.0f GB/s) | cuBLAS fp32 %.3f ms | cuBLAS tf32 %.3f ms"
              % (prec, M, K, N, t_tc, flops / t_tc / 1e9, byts / t_tc / 1e6, t_fp32, t_tf32), flush=True)
    if "--bwd" in sys.argv:
        for (M, K, N) in [(4096 * 256, 64, 416), (512 * 703, 192, 384), (65536, 1024, 1024)]:
            x = torch.randn(M, K, device=dev)
            w = torch.randn(N, K, device=dev) / K ** 0.5
            gy = torch.randn(M, N, device=dev)
            gw = torch.zeros(N, K, device=dev)
            t_dx = timeit(lambda: ops.linear_bwd(x, w, gy, need_weight=False, precision=prec))
            t_dw = timeit(lambda: ops.linear_bwd(x, w, gy, need_x=False, precision=prec, grad_weight=gw))
            torch.backends.cuda.matmul.allow_tf32 = True
            c_dx = timeit(lambda: gy @ w)
            c_dw = timeit(lambda: gy.t() @ x)
            torch.backends.cuda.matmul.allow_tf32 = False
            flops = 2.0 * M * N * K
            print("time bwd %-7s M=%d K=%d N=%d: grad_x %.3f ms (%.1f TFLOP/s) cuBLAS tf32 %.3f | grad_W %.3f ms (%.1f TFLOP/s) cuBLAS tf32 %.3f"
                  % (prec, M, K, N, t_dx, flops / t_dx / 1e9, c_dx, t_dw, flops / t_dw / 1e9, c_dw), flush=True)
sys.exit(1 if bad else 0)
