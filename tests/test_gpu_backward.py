"""GPU parity of the backward kernels: gradients through the fused layers against torch autograd through
the CPU oracle (the reference trains by autograd through exactly these eager float64 ops,
general/train.py:148-152) for a random linear functional of (z_out, ldj).

Tolerance: |g - g_ref| <= 2e-4 |g_ref| + 2e-5 * max|g_ref| per tensor (fp32 kernels vs the oracle's float64
internals), parameter gradients (long reductions) 5e-4 relative to their max."""
import pytest
import torch

from conftest import assert_close
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu


def grads_close(a, b, what, rtol=2e-4, atol_rel=2e-5):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    scale = max(b.abs().max().item(), 1e-12)
    err = (a - b).abs()
    tol = rtol * b.abs() + atol_rel * scale
    bad = err > tol
    assert not bad.any(), "%s: %d/%d out of tolerance, worst err %.3e (scale %.3e)" % (
        what, int(bad.sum()), bad.numel(), (err - tol).max().item(), scale)


def leaf(t, cuda=False):
    t = t.clone().cuda() if cuda else t.clone()
    return t.requires_grad_(True)


def _mix_inputs(B, S, C, K, seed, padded, chess=False, flip=False, ratio=0.5, nn_std=0.7):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, S, C, generator=g) * 1.3
    nn_out = torch.randn(B, S, C * (2 + 3 * K), generator=g) * nn_std
    sf, msf = torch.randn(C, generator=g) * 0.4, torch.randn(C, K, generator=g) * 0.4
    if chess:
        mask = torch.tensor([1.0, 0.0]).view(2, 1)
    else:
        n_cond = int(C * ratio)
        mask = torch.zeros(1, C)
        mask[0, :n_cond] = 1.0
    if flip:
        mask = 1 - mask
    pad = None
    if padded:
        lens = torch.randint(max(1, S // 2), S + 1, (B,), generator=g)
        pad = (torch.arange(S)[None, :] < lens[:, None]).float()
    wz, wl = torch.randn(B, S, C, generator=g), torch.randn(B, generator=g)
    return z, nn_out, sf, msf, mask, pad, wz, wl


@pytest.mark.parametrize("B,S,C,K,padded,chess,flip,reg", [
    (3, 32, 16, 8, False, False, False, False), (4, 40, 16, 8, True, False, False, True), (5, 38, 6, 16, True, False, False, True),
    (3, 71, 2, 8, True, False, False, False), (4, 9, 1, 4, True, True, False, False), (3, 6, 4, 10, False, False, True, False),
    (2, 5, 5, 3, False, False, False, False), (64, 64, 16, 8, False, False, False, False)])
def test_mixcdf_backward_vs_oracle_autograd(B, S, C, K, padded, chess, flip, reg):
    from categoricalnf_b200 import functional as CF
    z, nn_out, sf, msf, mask, pad, wz, wl = _mix_inputs(B, S, C, K, seed=B + S + C + K, padded=padded, chess=chess, flip=flip)
    reg_max, reg_factor = (1.0, 1.5) if reg else (-1.0, 1.0)
    # oracle + autograd (CPU)
    zo, no, so, mo = leaf(z), leaf(nn_out), leaf(sf), leaf(msf)
    m = O.expand_mask(mask, z)
    out, ldj, _ = O.mixcdf_coupling(zo, no, m, K, so, mo, pad=pad.unsqueeze(-1) if pad is not None else None,
                                    reg_max=reg_max, reg_factor=reg_factor, training=True)
    ((out * wz).sum() + (ldj * wl).sum()).backward()
    # kernels (GPU)
    zg, ng, sg, mg = leaf(z, True), leaf(nn_out, True), leaf(sf, True), leaf(msf, True)
    mc, ms = (mask.flatten().tolist(), None) if mask.shape[0] == 1 else (None, mask.flatten().tolist())
    out_g, ldj_g, _ = CF.mixcdf(zg, ng, K, sg, mg, mask_c=mc, mask_s=ms, pad=pad.cuda() if pad is not None else None,
                                reg_max=reg_max, reg_factor=reg_factor, training=True)
    ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
    grads_close(zg.grad, zo.grad, "dL/dz")
    grads_close(ng.grad, no.grad, "dL/dnn_out")
    grads_close(sg.grad, so.grad, "dL/dscaling_factor", rtol=5e-4, atol_rel=5e-4)
    grads_close(mg.grad, mo.grad, "dL/dmixture_scaling_factor", rtol=5e-4, atol_rel=5e-4)


def test_mixcdf_backward_at_zero_scaling_factors():
    """Freshly built layers have scaling_factor = mixture_scaling_factor = 0, i.e. exp(.) sits exactly ON the bound of
    ``scaling_fac.clamp(min=1.0)`` (mixture_cdf_layer.py:158-162), where torch passes the gradient through the clamp.
    Found by tests/test_gpu_reference_training.py: the first optimiser step of every training run starts here."""
    from categoricalnf_b200 import functional as CF
    B, S, C, K = 4, 12, 6, 8
    z, nn_out, sf, msf, mask, pad, wz, wl = _mix_inputs(B, S, C, K, seed=91, padded=False)
    sf, msf = torch.zeros_like(sf), torch.zeros_like(msf)
    msf[0, :3] = torch.tensor([0.3, -0.3, 0.0])          # both sides of the bound next to it
    zo, no, so, mo = leaf(z), leaf(nn_out), leaf(sf), leaf(msf)
    out, ldj, _ = O.mixcdf_coupling(zo, no, O.expand_mask(mask, z), K, so, mo, training=True)
    ((out * wz).sum() + (ldj * wl).sum()).backward()
    zg, ng, sg, mg = leaf(z, True), leaf(nn_out, True), leaf(sf, True), leaf(msf, True)
    out_g, ldj_g, _ = CF.mixcdf(zg, ng, K, sg, mg, mask_c=mask.flatten().tolist(), training=True)
    ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
    assert so.grad.abs().max() > 1e-3 and mo.grad.abs().max() > 1e-3
    grads_close(ng.grad, no.grad, "dL/dnn_out")
    grads_close(sg.grad, so.grad, "dL/dscaling_factor at 0", rtol=5e-4, atol_rel=5e-4)
    grads_close(mg.grad, mo.grad, "dL/dmixture_scaling_factor at 0", rtol=5e-4, atol_rel=5e-4)


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("zero_sf", [False, True])
def test_affine_backward(reverse, zero_sf):
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(3)
    B, S, C = 6, 17, 6
    z, nn_out, sf = torch.randn(B, S, C, generator=g), torch.randn(B, S, 2 * C, generator=g) * 0.6, torch.randn(C, generator=g) * 0.4
    if zero_sf:          # exp(0) = 1 is exactly the clamp bound (coupling_layer.py:82): torch passes the gradient there
        sf = torch.zeros_like(sf)
    wz, wl = torch.randn(B, S, C, generator=g), torch.randn(B, generator=g)
    mask = torch.zeros(1, C)
    mask[0, :3] = 1.0
    zo, no, so = leaf(z), leaf(nn_out), leaf(sf)
    out, ldj = O.affine_coupling(zo, no, mask.unsqueeze(0), so, reverse=reverse)
    ((out * wz).sum() + (ldj * wl).sum()).backward()
    zg, ng, sg = leaf(z, True), leaf(nn_out, True), leaf(sf, True)
    out_g, ldj_g = CF.affine_coupling(zg, ng, sg, mask_c=mask.flatten().tolist(), reverse=reverse)
    ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
    grads_close(zg.grad, zo.grad, "dL/dz")
    grads_close(ng.grad, no.grad, "dL/dnn_out")
    grads_close(sg.grad, so.grad, "dL/dscaling_factor", rtol=5e-4, atol_rel=5e-4)


@pytest.mark.parametrize("C", [6, 16, 12])      # 6: scalar kernel; 16 / 12: four channels per thread (actnorm_bwd4_kernel)
@pytest.mark.parametrize("reverse,padded,with_length", [(False, False, False), (False, True, False), (True, True, True), (False, False, True)])
def test_actnorm_backward(reverse, padded, with_length, C):
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(5)
    B, S = 7, 23
    z = torch.randn(B, S, C, generator=g)
    bias, scales = torch.randn(1, 1, C, generator=g) * 0.3, torch.randn(1, 1, C, generator=g) * 0.3
    wz, wl = torch.randn(B, S, C, generator=g), torch.randn(B, generator=g)
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    pad = (torch.arange(S)[None, :] < lens[:, None]).float() if padded else None
    length = lens if with_length else None
    zo, bo, so = leaf(z), leaf(bias), leaf(scales)
    out, ldj = O.actnorm(zo, bo, so, reverse=reverse, length=length, pad=pad.unsqueeze(-1) if padded else None)
    ((out * wz).sum() + (ldj * wl).sum()).backward()
    zg, bg, sg = leaf(z, True), leaf(bias, True), leaf(scales, True)
    ldj0 = torch.zeros(B, device="cuda")
    out_g, ldj_g = CF.actnorm(zg, bg, sg, ldj0, pad=pad.cuda() if padded else None,
                              length=length.float().cuda() if with_length else None, reverse=reverse)
    ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
    grads_close(zg.grad, zo.grad, "dL/dz")
    grads_close(bg.grad, bo.grad, "dL/dbias", rtol=5e-4, atol_rel=5e-4)
    grads_close(sg.grad, so.grad, "dL/dscales", rtol=5e-4, atol_rel=5e-4)


@pytest.mark.parametrize("reverse", [False, True])
def test_ext_actnorm_backward(reverse):
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(7)
    B, S, C = 9, 5, 4
    z, ext = torch.randn(B, S, C, generator=g), torch.randn(B, S, 2 * C, generator=g) * 0.7
    wz, wl = torch.randn(B, S, C, generator=g), torch.randn(B, generator=g)
    pad = (torch.rand(B, S, generator=g) > 0.3).float()
    zo, eo = leaf(z), leaf(ext)
    out, ldj = O.ext_actnorm(zo, eo[..., :C], eo[..., C:], reverse=reverse, pad=pad.unsqueeze(-1))
    ((out * wz).sum() + (ldj * wl).sum()).backward()
    zg, eg = leaf(z, True), leaf(ext, True)
    out_g, ldj_g = CF.ext_actnorm(zg, eg, torch.zeros(B, device="cuda"), pad=pad.cuda(), reverse=reverse)
    ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
    grads_close(zg.grad, zo.grad, "dL/dz")
    grads_close(eg.grad, eo.grad, "dL/dext")


@pytest.mark.parametrize("C,reverse,padded", [(16, False, False), (16, True, True), (16, False, True), (6, False, True), (2, True, True),
                                               (40, False, False)])
def test_invconv_backward(C, reverse, padded):
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(11 + C)
    B, S = 5, 29
    z = torch.randn(B, S, C, generator=g)
    w = torch.linalg.qr(torch.randn(C, C, generator=g))[0] + 0.05 * torch.randn(C, C, generator=g)
    sldj = torch.randn((), generator=g)
    wz, wl = torch.randn(B, S, C, generator=g), torch.randn(B, generator=g)
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    pad = (torch.arange(S)[None, :] < lens[:, None]).float() if padded else None
    zo, wo, so = leaf(z), leaf(w), leaf(sldj)
    out, ldj = O.invconv(zo, wo, so, reverse=reverse, length=lens if padded else None, pad=pad.unsqueeze(-1) if padded else None)
    ((out * wz).sum() + (ldj * wl).sum()).backward()
    zg, wg, sg = leaf(z, True), leaf(w, True), leaf(sldj, True)
    out_g, ldj_g = CF.invconv(zg, wg, sg, torch.zeros(B, device="cuda"), pad=pad.cuda() if padded else None,
                              length=lens.float().cuda() if padded else None, reverse=reverse)
    ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
    grads_close(zg.grad, zo.grad, "dL/dz")
    grads_close(wg.grad, wo.grad, "dL/dW", rtol=5e-4, atol_rel=5e-4)
    grads_close(sg.grad, so.grad, "dL/dsldj", rtol=5e-4, atol_rel=5e-4)


def test_logistic_logprob_backward():
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(13)
    x = torch.randn(4, 11, 6, generator=g) * 3
    w = torch.randn(4, 11, 6, generator=g)
    xo = leaf(x)
    (O.logistic_log_prob(xo) * w).sum().backward()
    xg = leaf(x, True)
    (CF.logistic_logprob(xg) * w.cuda()).sum().backward()
    grads_close(xg.grad, xo.grad, "dL/dx")


def test_tclinear_backward():
    from categoricalnf_b200.layers.networks import TCLinear
    g = torch.Generator().manual_seed(17)
    x = torch.randn(37, 9, 48, generator=g)
    ref = torch.nn.Linear(48, 72)
    mod = TCLinear(48, 72).cuda()
    mod.load_state_dict(ref.state_dict())
    w = torch.randn(37, 9, 72, generator=g)
    xo, xg = leaf(x), leaf(x, True)
    (ref(xo) * w).sum().backward()
    (mod(xg) * w.cuda()).sum().backward()
    grads_close(xg.grad, xo.grad, "dL/dx", rtol=1e-4, atol_rel=1e-5)
    grads_close(mod.weight.grad, ref.weight.grad, "dL/dW", rtol=1e-4, atol_rel=1e-5)
    grads_close(mod.bias.grad, ref.bias.grad, "dL/db", rtol=1e-4, atol_rel=1e-5)


def test_training_step_through_the_drop_in_flow():
    """One optimisation-style step through the module API (encoding -> [ActNorm, InvConv, MixtureCDFCoupling]
    x 2 -> prior): loss and every parameter gradient against autograd through the oracle composition."""
    import workload as W
    S, B = 24, 6
    prm = W.data_init_oracle(W.lm_params(seed=5, S=S, blocks=2), seed=5)
    tokens, u = W.lm_tokens(B, S, prm.V, seed=5), W.lm_noise(B, S, prm.D, seed=5)
    # ---- oracle with autograd over the raw parameters ------------------------------------------------
    names = ["bias", "scales", "l", "u", "log_s", "net_w", "net_b", "sf", "msf"]
    leaves = [{k: leaf(b[k]) for k in names} for b in prm.blocks]
    embed_w, pred_w, pred_b = leaf(prm.embed_w), leaf(prm.pred_w), leaf(prm.pred_b)
    table = torch.nn.functional.linear(embed_w, pred_w, pred_b)
    blocks = []
    for b, lv in zip(prm.blocks, leaves):
        w, sldj = O.invconv_weight(b["p"], lv["l"], lv["log_s"], lv["u"], b["sign_s"])
        blocks.append(dict(bias=lv["bias"].view(1, 1, -1), scales=lv["scales"].view(1, 1, -1), weight=w, sldj=sldj, mask=b["mask"],
                           K=prm.K, sf=lv["sf"], msf=lv["msf"],
                           nn_fn=(lambda zin, lv=lv: torch.nn.functional.linear(zin, lv["net_w"], lv["net_b"]))))
    z, ldj, lp = O.lm_flow_forward(tokens, u, dict(table=table, prior=prm.prior), blocks)
    loss_ref = ((-ldj - lp) / S).mean()
    loss_ref.backward()
    # ---- drop-in modules on the GPU ---------------------------------------------------------------------
    model, _ = W.build_lm_model(prm, torch.device("cuda", 0))
    model.train()
    from categoricalnf_b200 import functional as CF
    zg, ldj_g = model(tokens.cuda(), u_noise=u.cuda())
    lp_g = CF.logistic_logprob(zg).sum(dim=[1, 2])
    loss = ((-ldj_g - lp_g) / S).mean()
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item()) + 1e-5
    enc = model.flow_layers[0]
    grads_close(enc.embed_layer.weight.grad, embed_w.grad, "embed", rtol=1e-3, atol_rel=1e-3)
    grads_close(enc.flow_layers[0].pred_net.layer.weight.grad, pred_w.grad, "pred_net.weight", rtol=1e-3, atol_rel=1e-3)
    grads_close(enc.flow_layers[0].pred_net.layer.bias.grad, pred_b.grad, "pred_net.bias", rtol=1e-3, atol_rel=1e-3)
    for i, lv in enumerate(leaves):
        an, conv, mix = model.flow_layers[1 + 3 * i: 4 + 3 * i]
        grads_close(an.bias.grad.flatten(), lv["bias"].grad, "block %d actnorm bias" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(an.scales.grad.flatten(), lv["scales"].grad, "block %d actnorm scales" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(conv.l.grad, lv["l"].grad, "block %d conv l" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(conv.u.grad, lv["u"].grad, "block %d conv u" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(conv.log_s.grad, lv["log_s"].grad, "block %d conv log_s" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(mix.nn.lin.weight.grad, lv["net_w"].grad, "block %d net weight" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(mix.nn.lin.bias.grad, lv["net_b"].grad, "block %d net bias" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(mix.scaling_factor.grad, lv["sf"].grad, "block %d sf" % i, rtol=1e-3, atol_rel=1e-3)
        grads_close(mix.mixture_scaling_factor.grad, lv["msf"].grad, "block %d msf" % i, rtol=1e-3, atol_rel=1e-3)


@pytest.mark.parametrize("B,S,V,D,padded,beta", [(5, 17, 51, 16, False, 1.0), (4, 38, 9, 6, True, 0.7), (3, 50, 3, 2, True, 1.0),
                                                  (2, 12, 1, 2, False, 1.0), (6, 9, 100, 16, True, 0.5), (64, 256, 51, 16, False, 1.0)])
def test_categ_encode_backward_vs_oracle_autograd(B, S, V, D, padded, beta):
    """cnf_categ_encode_bwd: dL/dtable against autograd through the CPU oracle's all-class expansion
    (linear_encoding.py:71-92,153-174), with a loss that uses both outputs (z and ldj)."""
    from categoricalnf_b200 import functional as CF
    g = torch.Generator().manual_seed(B * S + V + D)
    table = torch.cat([torch.randn(V, D, generator=g) * 1.5, torch.randn(V, D, generator=g) * 0.4 - 0.5], dim=1)
    prior = torch.log_softmax(torch.randn(V, generator=g), 0)
    x = torch.randint(0, V, (B, S), generator=g)
    u = torch.rand(B * S, 1, D, generator=g)
    lens = torch.randint(1, S + 1, (B,), generator=g)
    pad = (torch.arange(S)[None, :] < lens[:, None]).float().unsqueeze(-1) if padded else None
    wz = torch.randn(B, S, D, generator=g)
    wl = torch.randn(B, generator=g)
    t_ref = table.double().requires_grad_(True)
    z_ref, ldj_ref, _ = O.categ_encode(x, u.double(), t_ref, prior.double(), beta=beta, pad=pad.double() if padded else None)
    ((z_ref * wz.double()).sum() + (ldj_ref * wl.double()).sum()).backward()
    t_gpu = table.cuda().requires_grad_(True)
    z, ldj, cpl = CF.categ_encode(x.cuda(), t_gpu, prior.cuda(), noise=u.cuda(), pad=pad.cuda() if padded else None, beta=beta)
    ((z * wz.cuda()).sum() + (ldj * wl.cuda()).sum()).backward()
    assert_close(z, z_ref, rtol=1e-4, atol=1e-5, what="z")
    grads_close(t_gpu.grad, t_ref.grad.float(), "dL/dtable", rtol=2e-3, atol_rel=1e-3)


@pytest.mark.parametrize("B,S,C,K,flip", [(5, 41, 16, 8, False), (4, 33, 16, 8, True), (3, 20, 16, 64, False), (2, 50, 16, 32, False),
                                          (3, 37, 8, 8, False), (600, 64, 16, 8, False)])
def test_mixcdf_compact_layout_matches_full(B, S, C, K, flip):
    """ABI v4 ``nn_compact``: the network output holds the transformed channels' records only.  Forward, inverse and
    backward on the compact tensor equal the full-layout calls (bit for bit forward, to rounding in the gradients), and
    dL/dnn_out comes back compact - nothing is written for conditioner channels."""
    from categoricalnf_b200 import functional as CF
    from categoricalnf_b200 import ops
    z, nn_out, sf, msf, mask, pad, wz, wl = _mix_inputs(B, S, C, K, seed=B + S + C + K, padded=True, flip=flip)
    mc = mask.flatten().tolist()
    PN = 2 + 3 * K
    tch = [c for c, m in enumerate(mc) if m == 0.0]
    r0, r1 = tch[0] * PN, (tch[-1] + 1) * PN
    nn_c = nn_out[..., r0:r1].contiguous()
    assert ops.mixcdf_path(z.cuda(), nn_c.cuda(), K, mask_c=mc, compact=True) in ("pipe", "gpipe")
    kw = dict(mask_c=mc, pad=pad.cuda(), scaling_factor=sf.cuda(), mixture_scaling_factor=msf.cuda())
    zf, lf, _ = ops.mixcdf(z.cuda(), nn_out.cuda(), K, **kw)
    zc, lc, _ = ops.mixcdf(z.cuda(), nn_c.cuda(), K, compact=True, **kw)
    assert torch.equal(zf, zc)
    assert_close(lc, lf, rtol=1e-6, atol=1e-5, what="ldj (atomics: order-dependent last bits)")
    zi, li, _ = ops.mixcdf(zf, nn_out.cuda(), K, reverse=True, **kw)
    zic, lic, _ = ops.mixcdf(zf, nn_c.cuda(), K, reverse=True, compact=True, **kw)
    assert torch.equal(zi, zic)
    assert_close(lic, li, rtol=1e-6, atol=1e-5, what="ldj inverse")
    # backward: full vs compact
    res = {}
    for compact in (False, True):
        zg, sg, mg = leaf(z, True), leaf(sf, True), leaf(msf, True)
        ng = leaf(nn_c if compact else nn_out, True)
        out_g, ldj_g, _ = CF.mixcdf(zg, ng, K, sg, mg, mask_c=mc, pad=pad.cuda(), reg_max=1.0, reg_factor=1.5, training=True,
                                    compact=compact)
        ((out_g * wz.cuda()).sum() + (ldj_g * wl.cuda()).sum()).backward()
        res[compact] = (zg.grad, ng.grad, sg.grad, mg.grad)
    assert res[True][1].shape == nn_c.shape
    grads_close(res[True][0], res[False][0], "dL/dz", rtol=1e-6, atol_rel=1e-7)
    grads_close(res[True][1], res[False][1][..., r0:r1], "dL/dnn_out (compact)", rtol=1e-6, atol_rel=1e-7)
    assert float(res[False][1][..., :r0].abs().sum() + res[False][1][..., r1:].abs().sum()) == 0.0
    grads_close(res[True][2], res[False][2], "dL/dsf", rtol=1e-4, atol_rel=1e-5)
    grads_close(res[True][3], res[False][3], "dL/dmsf", rtol=1e-4, atol_rel=1e-5)


@pytest.mark.parametrize("padded,flip,fused", [(False, False, True), (True, False, True), (True, True, True), (True, False, False)])
def test_compact_projection_in_training_matches_two_step_path(padded, flip, fused):
    """MixtureCDFCoupling in training mode with a network that ends in a Linear: the compact path (only the transformed
    channels' weight rows are multiplied) gives the same outputs and the same gradients - incl. exact zeros for the skipped
    weight rows - as the full network output through the two-step path."""
    import workload as W
    from categoricalnf_b200 import functional as CFm
    from categoricalnf_b200.layers.flows import MixtureCDFCoupling
    torch.manual_seed(0)
    D, K, B, S = 16, 8, 7, 50
    mask = torch.cat([torch.ones(D // 2), torch.zeros(D - D // 2)]).view(1, D)
    if flip:
        mask = 1 - mask
    pad = None
    if padded:
        lens = torch.randint(S // 2, S + 1, (B,))
        pad = (torch.arange(S)[None, :] < lens[:, None]).float().unsqueeze(-1).cuda()
    # fused = True: the per-position Linear's backward runs inside the transform's backward kernel (ABI v5 proj_weight)
    CFm.FUSE_LINEAR_BACKWARD = fused
    layer = MixtureCDFCoupling(c_in=D, mask=mask, model_func=lambda c_out: W.StandInNet(D, c_out), num_mixtures=K).cuda().train()
    with torch.no_grad():
        layer.scaling_factor.normal_(0, 0.3)
        layer.mixture_scaling_factor.normal_(0, 0.3)
    z = torch.randn(B, S, D, device="cuda")
    wz, wl = torch.randn(B, S, D, device="cuda"), torch.randn(B, device="cuda")
    out = {}
    for compact in (True, False):
        layer.compact_projection_in_training = compact
        layer.zero_grad()
        zin = z.clone().requires_grad_(True)
        zo, ldj, _ = layer(zin, channel_padding_mask=pad)
        ((zo * wz).sum() + (ldj * wl).sum()).backward()
        out[compact] = (zo.detach(), ldj.detach(), zin.grad, layer.nn.lin.weight.grad.clone(), layer.nn.lin.bias.grad.clone(),
                        layer.scaling_factor.grad.clone(), layer.mixture_scaling_factor.grad.clone())
    CFm.FUSE_LINEAR_BACKWARD = True
    names = ["z", "ldj", "dL/dz", "dL/dW", "dL/db", "dL/dsf", "dL/dmsf"]
    for a, b, n in zip(out[True], out[False], names):
        grads_close(a, b, n, rtol=2e-5, atol_rel=2e-6)
    pn = 2 + 3 * K
    cond_rows = slice((D // 2) * pn, None) if flip else slice(0, (D // 2) * pn)
    assert float(out[True][3][cond_rows].abs().sum()) == 0.0        # conditioner records: no gradient, as in the reference
