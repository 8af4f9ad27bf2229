"""Locate the element that makes cnf_mixcdf_inv return NaN in the K=16 reverse case of fused_check.py and
dump its inputs (gpurun_out/nan_case.pt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops
dev = torch.device("cuda", 0)
B, S, C, K, H = 3, 77, 8, 16, 64
g = torch.Generator().manual_seed(B * 131 + S * 7 + C + K + H)
PN = 2 + 3 * K
z = (torch.randn(B, S, C, generator=g) * 1.2).to(dev)
feats = torch.randn(B, S, H, generator=g).to(dev)
w = (torch.randn(C * PN, H, generator=g) * (0.5 / H ** 0.5)).to(dev)
b = (torch.randn(C * PN, generator=g) * 0.1).to(dev)
Ct = C // 2
mask_c = [1.0] * (C - Ct) + [0.0] * Ct
sf = (torch.randn(C, generator=g) * 0.3).to(dev)
msf = (torch.randn(C, K, generator=g) * 0.3).to(dev)
nn64 = (feats.double() @ w.double().t() + b.double()).float()
# one sample per position, one channel at a time
zz = z.reshape(B * S, 1, C)
nn = nn64.reshape(B * S, 1, C * PN)
for env in ("pipe", "generic"):
    zo, ldj, _ = ops.mixcdf(zz, nn, K, mask_c=mask_c, scaling_factor=sf, mixture_scaling_factor=msf, reverse=True)
    torch.cuda.synchronize()
    try:
        ops.check_status(dev, env)
    except Exception as e:
        print(env, e)
    idx = torch.isnan(ldj).nonzero().flatten().tolist()
    print("positions with NaN ldj:", idx, "nan in z:", torch.isnan(zo).sum().item())
    for i in idx[:4]:
        print("pos", i, "z", zz[i, 0].tolist())
        print("z_out", zo[i, 0].tolist())
        rec = nn[i, 0].reshape(C, PN)
        for c in range(C - Ct, C):
            print(" ch", c, "t,log_s", rec[c, :2].tolist(), "sf", sf[c].item())
            print("   log_pi", rec[c, 2:2 + K].tolist())
            print("   mu", rec[c, 2 + K:2 + 2 * K].tolist())
            print("   ls", rec[c, 2 + 2 * K:].tolist())
            print("   msf", msf[c].tolist())
    torch.save(dict(z=zz.cpu(), nn=nn.cpu(), sf=sf.cpu(), msf=msf.cpu(), idx=idx), "gpurun_out/nan_case.pt")
    break
