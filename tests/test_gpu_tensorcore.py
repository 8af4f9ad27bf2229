"""GPU parity of the tcgen05 kernels: the dense projection of the coupling networks (cnf_linear_fwd / _bwd) and
the fused final projection + mixture coupling (cnf_linear_mixcdf_fwd / _inv).

Tolerances: 3xTF32 projection |err| <= 6e-6 + 8e-8 K against a float64 product of O(1) outputs (the tensor core
truncates fp32 operands to TF32 and accumulates in fp32 with truncation: the dropped lo*lo term of a split done on
truncated high parts is a coherent ~2^-22 of sum |x w|, and the accumulation error grows with K); TF32 2e-2.  The fused kernel is held to
the flow's parity metric |a-b| <= 1e-4 |b| + 1e-5 (z), 1e-4 relative (ldj) against the CPU oracle applied to
a float64 projection, and against the golden fixtures of the reference (nn_out reproduced as features @ I)."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu


def dev(t):
    return t.cuda() if isinstance(t, torch.Tensor) else t


def _rand_linear(M, K, N, seed, bias=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    return x, w, b


@pytest.mark.parametrize("M,K,N", [(128, 32, 32), (100, 32, 64), (1000, 16, 416), (4096, 384, 208), (300, 100, 50),
                                   (777, 36, 418), (33, 1024, 1300), (1, 4, 8), (40000, 64, 512)])
@pytest.mark.parametrize("precision", ["3xtf32", "tf32"])
def test_linear_vs_float64(M, K, N, precision):
    from categoricalnf_b200 import ops
    x, w, b = _rand_linear(M, K, N, seed=M + K + N)
    y = ops.linear(dev(x), dev(w), dev(b), precision=precision)
    ref = x.double() @ w.double().t() + b.double()
    tol = (6e-6 + 8e-8 * K) if precision == "3xtf32" else 2e-2
    err = (y.double().cpu() - ref).abs().max().item()
    assert torch.isfinite(y).all()
    assert err <= tol, "max |err| %.3e > %.3e" % (err, tol)


def test_linear_gelu_nobias_batched_and_padded_k():
    from categoricalnf_b200 import ops
    x, w, _ = _rand_linear(6 * 50, 30, 96, seed=5, bias=False)      # K = 30 is zero-padded to 32 by the wrapper
    y = ops.linear(dev(x.reshape(6, 50, 30)), dev(w), None, activation="gelu")
    ref = torch.nn.functional.gelu(x.double() @ w.double().t()).reshape(6, 50, 96)
    assert y.shape == (6, 50, 96)
    assert_close(y, ref, rtol=1e-5, atol=5e-6, what="gelu(linear)")


def test_linear_rejects_cpu_and_bad_shapes():
    from categoricalnf_b200 import ops
    with pytest.raises(RuntimeError):
        ops.linear(torch.zeros(4, 8), torch.zeros(8, 8))
    with pytest.raises(ValueError):
        ops.linear(torch.zeros(4, 8, device="cuda"), torch.zeros(8, 12, device="cuda"))
    y = ops.linear(torch.zeros(0, 8, device="cuda"), torch.zeros(16, 8, device="cuda"))
    assert y.shape == (0, 16)


def test_tclinear_module_matches_nn_linear_state_dict():
    from categoricalnf_b200.layers.networks import TCLinear
    ref = torch.nn.Linear(48, 72)
    mod = TCLinear(48, 72).cuda()
    mod.load_state_dict(ref.state_dict())          # same parameter names as nn.Linear
    x = torch.randn(5, 17, 48)
    with torch.no_grad():
        assert_close(mod(x.cuda()), ref(x), rtol=1e-5, atol=5e-6, what="TCLinear")


@pytest.mark.parametrize("M,K,N", [(128, 32, 32), (100, 32, 64), (1000, 16, 416), (4096, 384, 208), (300, 100, 52),
                                   (777, 36, 420), (33, 1024, 1300), (1, 4, 8), (40000, 64, 512), (5000, 30, 50)])
@pytest.mark.parametrize("precision", ["3xtf32", "tf32"])
def test_linear_backward_vs_float64(M, K, N, precision):
    """cnf_linear_bwd: grad_x = gy W (MN-major B), grad_W = gy^T x (both MN-major, split over M, red.add),
    grad_b = column sums - against float64 products.  (5000, 30, 50) goes through the zero-padding wrapper."""
    from categoricalnf_b200 import ops
    x, w, _ = _rand_linear(M, K, N, seed=M + K + N + 1)
    gy = torch.randn(M, N, generator=torch.Generator().manual_seed(M + 7)) / N ** 0.5
    gx, gw, gb = ops.linear_bwd(dev(x), dev(w), dev(gy), need_bias=True, precision=precision)
    ref_x = gy.double() @ w.double()
    ref_w = gy.double().t() @ x.double()
    ref_b = gy.double().sum(dim=0)
    # error model: 3xTF32 keeps ~fp32 products, accumulation error grows with the reduction length
    tol_x = (6e-6 + 8e-8 * N) if precision == "3xtf32" else 2e-2
    scale_w = max(1.0, (M / N) ** 0.5)                # |grad_W| entries are O(sqrt(M/N))
    tol_w = ((6e-6 + 8e-8 * min(M, 4096)) if precision == "3xtf32" else 2e-2) * scale_w
    for got, ref, tol, what in ((gx, ref_x, tol_x, "grad_x"), (gw, ref_w, tol_w, "grad_weight"), (gb, ref_b, 1e-4 * scale_w, "grad_bias")):
        assert got.shape == ref.shape, what
        assert torch.isfinite(got).all(), what
        err = (got.double().cpu() - ref).abs().max().item()
        assert err <= tol, "%s: max |err| %.3e > %.3e" % (what, err, tol)


def test_linear_backward_accumulates_and_partial_outputs():
    from categoricalnf_b200 import ops
    x, w, _ = _rand_linear(3000, 64, 96, seed=11)
    gy = torch.randn(3000, 96, generator=torch.Generator().manual_seed(12)) * 0.1
    gw0 = torch.randn(96, 64, generator=torch.Generator().manual_seed(13))
    gb0 = torch.randn(96, generator=torch.Generator().manual_seed(14))
    gx, gw, gb = ops.linear_bwd(dev(x), dev(w), dev(gy), need_x=False, need_bias=True, grad_weight=dev(gw0.clone()),
                                grad_bias=dev(gb0.clone()))
    assert gx is None
    assert_close(gw, gw0.double() + gy.double().t() @ x.double(), rtol=1e-5, atol=2e-5, what="accumulated grad_weight")
    assert_close(gb, gb0.double() + gy.double().sum(0), rtol=1e-5, atol=2e-5, what="accumulated grad_bias")
    gx, gw, gb = ops.linear_bwd(dev(x), dev(w), dev(gy), need_weight=False)
    assert gw is None and gb is None
    assert_close(gx, gy.double() @ w.double(), rtol=1e-5, atol=5e-6, what="grad_x only")


def test_tclinear_autograd_matches_nn_linear():
    from categoricalnf_b200.layers.networks import TCLinear
    torch.manual_seed(3)
    ref = torch.nn.Linear(48, 72).double()
    mod = TCLinear(48, 72).cuda()
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    x = torch.randn(7, 33, 48)
    xr = x.double().requires_grad_(True)
    xg = x.cuda().requires_grad_(True)
    wgt = torch.randn(7, 33, 72)
    (ref(xr) * wgt.double()).sum().backward()
    (mod(xg) * wgt.cuda()).sum().backward()
    assert_close(xg.grad, xr.grad, rtol=1e-5, atol=5e-6, what="grad input")
    assert_close(mod.weight.grad, ref.weight.grad, rtol=1e-5, atol=2e-5, what="grad weight")
    assert_close(mod.bias.grad, ref.bias.grad, rtol=1e-5, atol=2e-5, what="grad bias")


# ---------------------------------------------------------------------------------------------------
# fused final projection + mixture coupling
# ---------------------------------------------------------------------------------------------------
def _fused_case(B, S, C, K, H, seed, first=True, padded=False, chess=False):
    g = torch.Generator().manual_seed(seed)
    PN = 2 + 3 * K
    z = torch.randn(B, S, C, generator=g) * 1.2
    feats = torch.randn(B, S, H, generator=g)
    w = torch.randn(C * PN, H, generator=g) * (0.5 / H ** 0.5)
    b = torch.randn(C * PN, generator=g) * 0.1
    sf, msf = torch.randn(C, generator=g) * 0.3, torch.randn(C, K, generator=g) * 0.3
    Ct = C // 2
    mask = torch.tensor(([1.0] * (C - Ct) + [0.0] * Ct) if first else ([0.0] * Ct + [1.0] * (C - Ct))).view(1, C)
    if chess:
        mask = torch.tensor([1.0, 0.0]).view(2, 1)
    pad = None
    if padded:
        lens = torch.randint(S // 2, S + 1, (B,), generator=g)
        pad = (torch.arange(S)[None, :] < lens[:, None]).float()
    return z, feats, w, b, sf, msf, mask, pad


@pytest.mark.parametrize("B,S,C,K,H,first,padded,chess", [
    (2, 64, 16, 8, 16, True, False, False), (8, 256, 16, 8, 16, True, True, False), (3, 100, 16, 8, 32, False, False, False),
    (5, 37, 8, 8, 64, True, True, False), (4, 64, 16, 8, 384, True, False, False), (4, 50, 8, 16, 128, True, False, False),
    (4, 50, 8, 4, 20, True, True, False), (6, 64, 16, 4, 48, False, False, False), (4, 64, 4, 8, 36, True, False, True),
    (150, 38, 16, 8, 64, True, True, False)])
def test_fused_projection_mixture_vs_oracle(B, S, C, K, H, first, padded, chess):
    from categoricalnf_b200 import ops
    z, feats, w, b, sf, msf, mask, pad = _fused_case(B, S, C, K, H, seed=B + S + C + K + H, first=first, padded=padded, chess=chess)
    nn_out = (feats.double() @ w.double().t() + b.double()).float()
    m = O.expand_mask(mask, z)
    pad3 = pad.unsqueeze(-1) if pad is not None else None
    zo, lo, ro = O.mixcdf_coupling(z, nn_out, m, K, sf, msf, pad=pad3, reg_max=3.0, reg_factor=1.5)
    mc, ms = (mask.flatten().tolist(), None) if mask.shape[0] == 1 else (None, mask.flatten().tolist())
    kw = dict(mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf), mixture_scaling_factor=dev(msf))
    assert ops.linear_mixcdf_fusable(dev(z), dev(feats), dev(w), K, mask_c=mc, mask_s=ms)
    zg, lg, rg = ops.linear_mixcdf(dev(z), dev(feats), dev(w), dev(b), K, reg_max=3.0, reg_factor=1.5, training=True,
                                   want_reg=True, **kw)
    ops.check_status(zg.device)
    assert_close(zg, zo, what="z")
    assert_close(lg, lo, rtol=1e-4, atol=2e-4, what="ldj")
    assert_close(rg, ro, rtol=1e-4, atol=2e-4, what="reg_ldj")
    # inverse of the forward output returns the input; ldj antisymmetric
    zr, lr, _ = O.mixcdf_coupling(zo, nn_out, m, K, sf, msf, pad=pad3, reverse=True)
    zgr, lgr, _ = ops.linear_mixcdf(dev(zo), dev(feats), dev(w), dev(b), K, reverse=True, **kw)
    ops.check_status(zg.device)
    assert_close(zgr, zr, what="z (inverse)")
    assert_close(lgr, lr, rtol=1e-4, atol=2e-4, what="ldj (inverse)")


@pytest.mark.parametrize("name", ["mixcdf_lm_small", "mixcdf_lm_padded_sf", "mixcdf_stress"])
def test_fused_reproduces_reference_golden(name):
    """The reference's golden (z, nn_out) -> (z_fwd, ldj): nn_out is fed as features through an identity
    projection restricted to what the kernel needs (features = nn_out, weight = I), so the fused kernel
    must reproduce the reference output itself."""
    from categoricalnf_b200 import ops
    g = load_golden(name)
    C, K = g.z.shape[-1], g.K
    PN = 2 + 3 * K
    S = g.z.shape[1]
    pad = (torch.arange(S).view(1, S) < g.length.view(-1, 1)).float() if g.get("padded", 0) else None
    eye = torch.eye(C * PN)
    kw = dict(mask_c=g.mask.flatten().tolist(), pad=dev(pad), scaling_factor=dev(g.sf), mixture_scaling_factor=dev(g.msf))
    z, ldj, _ = ops.linear_mixcdf(dev(g.z), dev(g.nn_out), dev(eye), None, K, **kw)
    ops.check_status(z.device)
    assert_close(z, g.z_fwd, what="z_fwd")
    assert_close(ldj, g.ldj_fwd, rtol=1e-4, atol=2e-4, what="ldj_fwd")
    z, ldj, _ = ops.linear_mixcdf(dev(g.z_lat), dev(g.nn_out), dev(eye), None, K, reverse=True, **kw)
    assert_close(z, g.z_smp, what="z_smp")
    assert_close(ldj, g.ldj_smp, rtol=1e-4, atol=2e-4, what="ldj_smp")


def test_fused_full_size_properties():
    """BASELINE size (B 4096 x S 256 x d 16, K 8): forward -> inverse round trip and batch-split linearity."""
    from categoricalnf_b200 import ops
    B, S, C, K, H = 4096, 256, 16, 8, 16
    gen = torch.Generator(device="cuda").manual_seed(3)
    z = torch.randn(B, S, C, device="cuda", generator=gen)
    feats = torch.randn(B, S, H, device="cuda", generator=gen)
    w = torch.randn(C * (2 + 3 * K), H, device="cuda", generator=gen) * 0.125
    b = torch.randn(C * (2 + 3 * K), device="cuda", generator=gen) * 0.1
    mc = [1.0] * 8 + [0.0] * 8
    zf, lf, _ = ops.linear_mixcdf(z, feats, w, b, K, mask_c=mc)
    zr, lr, _ = ops.linear_mixcdf(zf, feats, w, b, K, mask_c=mc, reverse=True)
    ops.check_status(z.device)
    assert (zr - z).abs().max().item() < 2e-3
    assert ((lf + lr).abs() / (lf.abs() + 1.0)).max().item() < 1e-4
    z2, l2, _ = ops.linear_mixcdf(z[100:200], feats[100:200], w, b, K, mask_c=mc)
    assert torch.equal(z2, zf[100:200])
    assert_close(l2, lf[100:200], rtol=1e-5, atol=1e-4, what="ldj of a batch slice")


@pytest.mark.parametrize("B,S,C,K,H,padded", [(4, 64, 16, 8, 16, False), (3, 50, 8, 8, 32, True), (5, 33, 16, 4, 64, True),
                                              (2, 300, 8, 16, 16, False)])
def test_fused_next_block_epilogue_and_masked_output(B, S, C, K, H, padded):
    """mix.next_* epilogue of the fused projection kernel (+ the next coupling's masked input) against the same
    layers run one kernel at a time."""
    from categoricalnf_b200 import ops
    z, feats, w, b, sf, msf, mask, pad = _fused_case(B, S, C, K, H, seed=7 * B + S + C + K + H, padded=padded)
    g = torch.Generator().manual_seed(99)
    nb, ns = torch.randn(C, generator=g) * 0.2, torch.randn(C, generator=g) * 0.2
    nw = torch.linalg.qr(torch.randn(C, C, generator=g))[0].contiguous()
    nmask = torch.cat([torch.zeros(C // 2), torch.ones(C - C // 2)])
    kw = dict(mask_c=mask.flatten().tolist(), pad=dev(pad), scaling_factor=dev(sf), mixture_scaling_factor=dev(msf))
    z1, l1, _ = ops.linear_mixcdf(dev(z), dev(feats), dev(w), dev(b), K, **kw)
    z1, _ = ops.actnorm(z1, dev(nb), dev(ns), None, pad=dev(pad))
    sldj = torch.zeros(1, device="cuda")
    z1, _ = ops.invconv_apply(z1, dev(nw), sldj, None, pad=dev(pad))
    z2, l2, _, zm = ops.linear_mixcdf(dev(z), dev(feats), dev(w), dev(b), K, fuse_next=(dev(nb), dev(ns), dev(nw)),
                                      next_mask=dev(nmask), **kw)
    ops.check_status(z2.device)
    assert_close(z2, z1, rtol=1e-5, atol=2e-6, what="z after the fused next block")
    assert_close(l2, l1, rtol=1e-6, atol=1e-5, what="ldj")
    assert torch.equal(zm, z2 * dev(nmask))


@pytest.mark.parametrize("M,K,N", [(1000, 16, 416), (4096, 384, 208), (777, 36, 418), (20480, 384, 2316)])
def test_linear_presplit_weight(M, K, N):
    """nn.Parameter weights take the pre-split path (high / low parts cached per parameter version, the kernel splits only
    the activations): same result as the in-kernel split within the 3xTF32 error, and the cache follows in-place updates."""
    from categoricalnf_b200 import ops
    x, w, b = _rand_linear(M, K, N, seed=M + K + N + 3)
    wp = torch.nn.Parameter(w.cuda())
    ref = x.double() @ w.double().t() + b.double()
    y = ops.linear(x.cuda(), wp, b.cuda())
    tol = 4e-6 + 8e-8 * K
    err = (y.double().cpu() - ref).abs().max().item()
    assert err <= tol, "pre-split: max |err| %.3e > %.3e" % (err, tol)
    assert ops.weight_split(wp) is not None and ops.weight_split(w.cuda()) is None
    with torch.no_grad():
        wp.mul_(1.5)                                           # version bump -> split recomputed
    y2 = ops.linear(x.cuda(), wp, b.cuda())
    ref2 = x.double() @ (1.5 * w.double()).t() + b.double()
    err2 = (y2.double().cpu() - ref2).abs().max().item()
    assert err2 <= 1.5 * tol, "after update: max |err| %.3e" % err2
    # the reference's RAdam writes weights through p.data (general/radam.py:82): no version bump - the optimiser post-step
    # hook (ops.param_epoch) is what invalidates the cached split on the evaluation path
    with torch.no_grad():
        ops.linear(x.cuda(), wp, b.cuda())                     # cached split of the current weight
        version = wp._version
        wp.data.mul_(2.0)
        assert wp._version == version
        torch.optim.SGD([wp], lr=0.0).step()
        y3 = ops.linear(x.cuda(), wp, b.cuda())
    ref3 = x.double() @ (3.0 * w.double()).t() + b.double()
    err3 = (y3.double().cpu() - ref3).abs().max().item()
    assert err3 <= 3.0 * tol, "after a p.data update + optimizer step: max |err| %.3e (stale cached split?)" % err3
