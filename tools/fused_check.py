"""GPU check of cnf_linear_mixcdf_fwd/_inv (final projection + mixture coupling in one kernel) against
the two-kernel path (cnf_linear_fwd 3xTF32 -> cnf_mixcdf_*) and a float64 projection, then timings.
    python tools/fused_check.py [--time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops

dev = torch.device("cuda", 0)
bad = 0


def case(B, S, C, K, H, ratio_first=True, pad=False, sfs=True, reverse=False, precision="3xtf32", chess=False, reg=False):
    global bad
    g = torch.Generator().manual_seed(B * 131 + S * 7 + C + K + H)
    PN = 2 + 3 * K
    z = (torch.randn(B, S, C, generator=g) * 1.2).to(dev)
    feats = torch.randn(B, S, H, generator=g).to(dev)
    w = (torch.randn(C * PN, H, generator=g) * (0.5 / H ** 0.5)).to(dev)
    b = (torch.randn(C * PN, generator=g) * 0.1).to(dev)
    Ct = C // 2
    mask_c = ([1.0] * (C - Ct) + [0.0] * Ct) if ratio_first else ([0.0] * Ct + [1.0] * (C - Ct))
    mask_s = None
    if chess:
        mask_c, mask_s = None, [1.0, 0.0]
    sf = (torch.randn(C, generator=g) * 0.3).to(dev) if sfs else None
    msf = (torch.randn(C, K, generator=g) * 0.3).to(dev) if sfs else None
    padm = None
    if pad:
        lens = torch.randint(S // 2, S + 1, (B,), generator=g)
        padm = (torch.arange(S)[None, :] < lens[:, None]).float().to(dev)
    kw = dict(mask_c=mask_c, mask_s=mask_s, pad=padm, scaling_factor=sf, mixture_scaling_factor=msf, reverse=reverse,
              reg_max=(2.0 if reg else -1.0), reg_factor=0.5, training=reg, want_reg=True)
    if not ops.linear_mixcdf_fusable(z, feats, w, K, mask_c=mask_c, mask_s=mask_s):
        print("B=%d S=%d C=%d K=%d H=%d: not fusable" % (B, S, C, K, H)); bad += 1; return
    nn64 = (feats.double() @ w.double().t() + b.double()).float()
    def status(where):
        try:
            ops.check_status(dev, where)
            return ""
        except (AssertionError, RuntimeError) as e:
            return " [%s: %s]" % (where, e)
    z_ref, ldj_ref, reg_ref = ops.mixcdf(z, nn64, K, **kw)
    note = status("two-kernel")
    z_f, ldj_f, reg_f = ops.linear_mixcdf(z, feats, w, b, K, precision=precision, **kw)
    torch.cuda.synchronize()
    note += status("fused")
    if note:
        note += " nan(ref z,ldj)=%d,%d nan(fused z,ldj)=%d,%d" % (torch.isnan(z_ref).sum(), torch.isnan(ldj_ref).sum(),
                                                                 torch.isnan(z_f).sum(), torch.isnan(ldj_f).sum())
    tol = 1.0 if precision == "3xtf32" else 300.0
    dz = ((z_f - z_ref).abs() / (1e-4 * z_ref.abs() + 1e-5)).max().item()
    dl = ((ldj_f - ldj_ref).abs() / (1e-4 * ldj_ref.abs() + 2e-4)).max().item()
    dr = (reg_f - reg_ref).abs().max().item() if reg else 0.0
    ok = dz <= tol and dl <= tol and dr <= 1e-3 * tol and bool(torch.isfinite(z_f).all())
    bad += (not ok)
    print("B=%-5d S=%-4d C=%-3d K=%-3d H=%-4d first=%d pad=%d sf=%d rev=%d chess=%d reg=%d %-6s z/tol=%.3f ldj/tol=%.3f reg=%.1e %s"
          % (B, S, C, K, H, ratio_first, pad, sfs, reverse, chess, reg, precision, dz, dl, dr, ("ok" if ok else "FAIL") + note), flush=True)


case(2, 64, 16, 8, 16)
case(8, 256, 16, 8, 16, pad=True)
case(3, 100, 16, 8, 32, ratio_first=False)
case(5, 37, 8, 8, 64, pad=True)
case(4, 256, 16, 8, 384)
case(4, 50, 8, 16, 128)
case(4, 50, 8, 4, 20, pad=True)
case(6, 64, 16, 4, 48, reg=True)
case(4, 64, 4, 8, 36, chess=True)
case(8, 256, 16, 8, 16, reverse=True)
case(3, 77, 8, 16, 64, reverse=True, pad=True)
case(64, 256, 16, 8, 16, precision="tf32")
case(700, 38, 16, 8, 64, pad=True)

if "--time" in sys.argv:
    def timeit(fn, reps=10):
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
    B, S, C, K = 4096, 256, 16, 8
    PN = 2 + 3 * K
    for H in (16, 64, 384):
        z = torch.randn(B, S, C, device=dev)
        feats = torch.randn(B, S, H, device=dev)
        w = torch.randn(C * PN, H, device=dev) * (0.5 / H ** 0.5)
        b = torch.randn(C * PN, device=dev) * 0.1
        mask_c = [1.0] * 8 + [0.0] * 8
        for prec in ("3xtf32", "tf32"):
            t_f = timeit(lambda: ops.linear_mixcdf(z, feats, w, b, K, mask_c=mask_c, precision=prec))
            t_l = timeit(lambda: ops.linear(feats, w, b, precision=prec))
            nn_out = ops.linear(feats, w, b, precision=prec)
            t_m = timeit(lambda: ops.mixcdf(z, nn_out, K, mask_c=mask_c))
            del nn_out
            print("time H=%d %-6s: fused %.3f ms | linear %.3f + mixcdf %.3f = %.3f ms" % (H, prec, t_f, t_l, t_m, t_l + t_m), flush=True)
        t_i = timeit(lambda: ops.linear_mixcdf(z, feats, w, b, K, mask_c=mask_c, reverse=True), reps=5)
        print("time H=%d inverse fused %.3f ms" % (H, t_i), flush=True)
sys.exit(1 if bad else 0)
