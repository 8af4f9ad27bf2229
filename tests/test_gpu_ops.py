"""GPU parity: every C-ABI kernel against the golden outputs of the unmodified reference
(tests/golden) and against the CPU oracle on fresh seeded inputs.

Tolerance (BASELINE.json north_star / SURVEY.md 8d): |a-b| <= 1e-4 |b| + 1e-5 on z,
1e-4 relative on ldj."""
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu

MIXCDF_CASES = ["mixcdf_selftest", "mixcdf_lm_small", "mixcdf_lm_padded_sf", "mixcdf_stress", "mixcdf_mol_nodes",
                "mixcdf_mol_edges", "mixcdf_chess", "mixcdf_chess_flip", "mixcdf_flip_k10", "mixcdf_ratio_k3"]


def dev(t):
    return t.cuda() if isinstance(t, torch.Tensor) else t


def split_mask(mask):
    """reference mask tensor -> (mask_c list | None, mask_s list | None)"""
    if mask.shape[0] == 1:
        return mask.flatten().tolist(), None
    return None, mask.flatten().tolist()


def pad_of(g):
    if not g.get("padded", 0):
        return None
    S = g.z.shape[1]
    return (torch.arange(S).view(1, S) < g.length.view(-1, 1)).float()


def ldj_close(a, b, what):
    assert_close(a, b, rtol=1e-4, atol=2e-4, what=what)


@pytest.mark.parametrize("name", MIXCDF_CASES)
def test_mixcdf_forward_golden(name):
    from categoricalnf_b200 import ops
    g = load_golden(name)
    mc, ms = split_mask(g.mask)
    pad = pad_of(g)
    z, ldj, reg = ops.mixcdf(dev(g.z), dev(g.nn_out), g.K, mask_c=mc, mask_s=ms, pad=dev(pad),
                             scaling_factor=dev(g.sf), mixture_scaling_factor=dev(g.msf),
                             reg_max=g.get("reg_max", -1.0), reg_factor=g.get("reg_factor", 1.0),
                             training=bool(g.get("training", 1)), want_reg=True)
    ops.check_status(z.device)
    assert_close(z, g.z_fwd, what="z_fwd")
    ldj_close(ldj, g.ldj_fwd, "ldj_fwd")
    if "reg_ldj" in g:
        ldj_close(reg, g.reg_ldj, "reg_ldj")


@pytest.mark.parametrize("name", MIXCDF_CASES)
def test_mixcdf_inverse_golden(name):
    from categoricalnf_b200 import ops
    g = load_golden(name)
    mc, ms = split_mask(g.mask)
    pad = pad_of(g)
    kw = dict(mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(g.sf), mixture_scaling_factor=dev(g.msf),
              reverse=True)
    nn_rev = g.get("nn_out_rev", g.nn_out)
    z, ldj, _ = ops.mixcdf(dev(g.z_fwd), dev(nn_rev), g.K, **kw)
    ops.check_status(z.device)
    assert_close(z, g.z_rev, what="z_rev")
    ldj_close(ldj, g.ldj_rev, "ldj_rev")
    if "z_lat" in g:
        z, ldj, _ = ops.mixcdf(dev(g.z_lat), dev(g.nn_out), g.K, **kw)
        assert_close(z, g.z_smp, what="z_smp")
        ldj_close(ldj, g.ldj_smp, "ldj_smp")


@pytest.mark.parametrize("per_position", [False, True])
def test_mixcdf_inverse_underflow_regression(per_position):
    """Inverse whose bisection midpoint lands where 1-F and the density underflow in float32: a Newton
    step formed from a flushed density must be rejected (it produced x = -inf once)."""
    from categoricalnf_b200 import ops
    g = load_golden("mixcdf_inv_underflow")
    mc, _ = split_mask(g.mask)
    z_lat, nn_out, z_ref = g.z_lat, g.nn_out, g.z_smp
    if per_position:   # [B*S, 1, C]: generic tile shapes as well
        z_lat, nn_out, z_ref = (t.reshape(-1, 1, t.shape[-1]) for t in (z_lat, nn_out, z_ref))
    z, ldj, _ = ops.mixcdf(dev(z_lat), dev(nn_out), g.K, mask_c=mc, scaling_factor=dev(g.sf),
                           mixture_scaling_factor=dev(g.msf), reverse=True)
    ops.check_status(z.device)
    assert_close(z, z_ref, what="z_smp")
    if not per_position:
        ldj_close(ldj, g.ldj_smp, "ldj_smp")


def test_mixcdf_tails_golden():
    """Deep left tails (CDF down to exp(-850), the 1e-22 clamps) and 1-CDF down to 1e-11."""
    from categoricalnf_b200 import ops
    g = load_golden("mixcdf_tails")
    mc, ms = split_mask(g.mask)
    z, ldj, _ = ops.mixcdf(dev(g.z), dev(g.nn_out), g.K, mask_c=mc, scaling_factor=dev(g.sf),
                           mixture_scaling_factor=dev(g.msf))
    assert_close(z, g.z_fwd, what="z")
    ldj_close(ldj, g.ldj_fwd, "ldj")


def test_mixcdf_right_tail_sanity():
    """CDF > 1 - 1e-13: the reference output is float64 round-off noise of `1 - exp(log_cdf)`
    (log(1-F) jumps between ~-36.7 and the -50.66 clamp), so only elements outside that band are
    compared; inside it the result must be finite and within the reference's attainable range."""
    from categoricalnf_b200 import ops
    g = load_golden("mixcdf_right_tail")
    mc, ms = split_mask(g.mask)
    z, ldj, _ = ops.mixcdf(dev(g.z), dev(g.nn_out), g.K, mask_c=mc, scaling_factor=dev(g.sf),
                           mixture_scaling_factor=dev(g.msf))
    z = z.cpu()
    m = O.expand_mask(g.mask, g.z)
    p = O.mixt_params(g.nn_out, m, g.K, g.sf, g.msf)
    surv = 1 - O._mix_log_cdf(g.z.double(), *p[2:]).exp()
    unstable = (surv < 1e-12) & (m == 0)
    assert unstable.sum() >= 1
    assert_close(z[~unstable], g.z_fwd[~unstable], what="z outside the unstable band")
    assert torch.isfinite(z[unstable]).all() and torch.isfinite(ldj).all()
    y = z[unstable].double() / p[1].exp()[unstable] - p[0][unstable]
    assert ((y > 27.0) & (y < 50.7)).all()
    # per-sample ldj: each unstable element may differ by at most log(1e-12) - log(1e-22) ~ 23
    assert ((ldj.cpu() - g.ldj_fwd).abs() <= 23.1 * unstable.sum(dim=[1, 2]) + 2e-3 * g.ldj_fwd.abs()).all()


def test_mixcdf_accumulate_and_autoregressive():
    from categoricalnf_b200 import ops
    g = load_golden("autoregressive_mixcdf")
    ldj = dev(g.ldj_in.clone())
    z, ldj2, _ = ops.mixcdf(dev(g.z), dev(g.nn_out), g.K, scaling_factor=dev(g.sf),
                            mixture_scaling_factor=dev(g.msf), ldj=ldj)
    assert ldj2.data_ptr() == ldj.data_ptr()
    assert_close(z * dev(g.pad), g.z_out, what="z")
    ldj_close(ldj2, g.ldj_out, "ldj")


@pytest.mark.parametrize("B,S,C,K,seed", [(7, 33, 16, 8, 0), (5, 19, 6, 16, 1), (3, 130, 2, 8, 2), (2, 9, 8, 4, 3),
                                          (2, 5, 3, 51, 4), (300, 1, 4, 8, 5), (1, 1, 1, 1, 6), (2, 40, 16, 64, 7)])
def test_mixcdf_vs_oracle_random(B, S, C, K, seed):
    from categoricalnf_b200 import ops
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, S, C, generator=g) * 1.3
    nn_out = torch.randn(B, S, C * (2 + 3 * K), generator=g) * 0.8
    sf, msf = torch.randn(C, generator=g) * 0.3, torch.randn(C, K, generator=g) * 0.3
    mask = O.channel_mask(C) if C > 1 else O.chess_mask()
    length = torch.randint(1, S + 1, (B,), generator=g)
    pad = (torch.arange(S).view(1, S) < length.view(-1, 1)).float().unsqueeze(-1)
    m = O.expand_mask(mask, z)
    zo, lo, ro = O.mixcdf_coupling(z, nn_out, m, K, sf, msf, pad=pad, reg_max=3.0, reg_factor=1.5)
    mc, ms = split_mask(mask)
    zg, lg, rg = ops.mixcdf(dev(z), dev(nn_out), K, mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf),
                            mixture_scaling_factor=dev(msf), reg_max=3.0, reg_factor=1.5, training=True, want_reg=True)
    assert_close(zg, zo, what="z_fwd")
    ldj_close(lg, lo, "ldj_fwd")
    ldj_close(rg, ro, "reg")
    zr, lr, _ = O.mixcdf_coupling(zo, nn_out, m, K, sf, msf, pad=pad, reverse=True)
    zgr, lgr, _ = ops.mixcdf(dev(zo), dev(nn_out), K, mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf),
                             mixture_scaling_factor=dev(msf), reverse=True)
    assert_close(zgr, zr, what="z_rev")
    ldj_close(lgr, lr, "ldj_rev")


def test_mixcdf_empty_and_status():
    from categoricalnf_b200 import ops
    z = torch.zeros(0, 5, 4, device="cuda")
    out, ldj, _ = ops.mixcdf(z, torch.zeros(0, 5, 4 * 26, device="cuda"), 8, mask_c=[1, 1, 0, 0])
    assert out.shape == (0, 5, 4) and ldj.shape == (0,)
    z = torch.randn(2, 3, 4, device="cuda")
    nn = torch.randn(2, 3, 4 * 26, device="cuda")
    nn[0, 0, 3 * 26 + 5] = float("nan")
    ops.mixcdf(z, nn, 8, mask_c=[1, 1, 0, 0])
    with pytest.raises(AssertionError):
        ops.check_status(z.device)
    ops.check_status(z.device)  # cleared
    with pytest.raises(RuntimeError):
        ops.mixcdf(z.cpu(), nn.cpu(), 8, mask_c=[1, 1, 0, 0])


@pytest.mark.parametrize("name", ["affine_coupling", "affine_coupling_tokens"])
def test_affine_golden(name):
    from categoricalnf_b200 import ops
    g = load_golden(name)
    mc, ms = split_mask(g.mask)
    ldj = dev(g.ldj_in.clone())
    z, ldj = ops.affine_coupling(dev(g.z), dev(g.nn_out), ldj, mask_c=mc, mask_s=ms, scaling_factor=dev(g.sf))
    assert_close(z, g.z_fwd, what="z_fwd")
    ldj_close(ldj, g.ldj_fwd, "ldj_fwd")
    zr, lr = ops.affine_coupling(dev(g.z_fwd), dev(g.nn_out), dev(g.ldj_fwd.clone()), mask_c=mc, mask_s=ms,
                                 scaling_factor=dev(g.sf), reverse=True)
    assert_close(zr, g.z_rev, what="z_rev")
    ldj_close(lr, g.ldj_rev, "ldj_rev")


def test_actnorm_golden():
    from categoricalnf_b200 import ops
    g = load_golden("actnorm")
    for tag, kw in (("plain", {}), ("len", dict(length=dev(g.length.float()), pad=dev(g.pad))),
                    ("padonly", dict(pad=dev(g.pad)))):
        ldj = dev(g["ldj_in_" + tag].clone())
        z, ldj2 = ops.actnorm(dev(g.z), dev(g.bias), dev(g.scales), ldj, **kw)
        assert ldj2.data_ptr() == ldj.data_ptr()
        assert_close(z, g["z_fwd_" + tag], what="z_fwd " + tag)
        ldj_close(ldj, g["ldj_fwd_" + tag], "ldj_fwd " + tag)
        zr, lr = ops.actnorm(dev(g["z_fwd_" + tag]), dev(g.bias), dev(g.scales), dev(g["ldj_fwd_" + tag].clone()),
                             reverse=True, **kw)
        assert_close(zr, g["z_rev_" + tag], what="z_rev " + tag)
        ldj_close(lr, g["ldj_rev_" + tag], "ldj_rev " + tag)
    b, s = ops.actnorm_data_init(dev(g.z), dev(g.pad))
    assert_close(b, g.init_bias_pad.flatten(), what="init bias")
    assert_close(s, g.init_scales_pad.flatten(), what="init scales")
    b, s = ops.actnorm_data_init(dev(g.z))
    assert_close(b, g.init_bias.flatten(), what="init bias nopad")
    assert_close(s, g.init_scales.flatten(), what="init scales nopad")


def test_ext_actnorm_golden():
    from categoricalnf_b200 import ops
    g = load_golden("ext_actnorm")
    ext = torch.nn.functional.linear(g.ext, g.weight, g.bias)
    z, ldj = ops.ext_actnorm(dev(g.z), dev(ext), dev(g.ldj_in.clone()), pad=dev(g.pad))
    assert_close(z, g.z_fwd, what="z_fwd")
    ldj_close(ldj, g.ldj_fwd, "ldj_fwd")
    zr, lr = ops.ext_actnorm(dev(g.z_fwd), dev(ext), dev(g.ldj_fwd.clone()), pad=dev(g.pad), reverse=True)
    assert_close(zr, g.z_rev, what="z_rev")
    ldj_close(lr, g.ldj_rev, "ldj_rev")
    z, ldj = ops.ext_actnorm(dev(g.z), dev(ext), dev(g.ldj_in.clone()))
    assert_close(z, g.z_fwd_nopad, what="z nopad")
    ldj_close(ldj, g.ldj_fwd_nopad, "ldj nopad")


@pytest.mark.parametrize("C", [2, 6, 16])
def test_invconv_golden(C):
    from categoricalnf_b200 import ops
    g = load_golden("invconv")
    t = "_c%d" % C
    w, w_inv, sldj = ops.invconv_build(p=dev(g["p" + t]), l=dev(g["l" + t]), u=dev(g["u" + t]),
                                       log_s=dev(g["log_s" + t]), sign_s=dev(g["sign_s" + t]))
    assert_close(w, g["w" + t], what="W")
    assert_close(w_inv, g["w_inv" + t], what="W_inv")
    assert_close(sldj, torch.tensor([g["sldj" + t]]), what="sldj")
    # direct parametrisation: log|det W| by in-kernel LU must equal sum(log_s)
    w2, w2_inv, sldj2 = ops.invconv_build(weight=dev(g["w" + t]))
    assert_close(sldj2, torch.tensor([g["sldj" + t]]), rtol=1e-4, atol=1e-5, what="slogdet")
    assert_close(w2_inv, g["w_inv" + t], what="W_inv direct")
    kw = dict(length=dev(g["length" + t].float()), pad=dev(g["pad" + t]))
    z, ldj = ops.invconv_apply(dev(g["z" + t]), w, sldj, dev(g["ldj_in" + t].clone()), **kw)
    assert_close(z, g["z_fwd" + t], what="z_fwd")
    ldj_close(ldj, g["ldj_fwd" + t], "ldj_fwd")
    zr, lr = ops.invconv_apply(dev(g["z_fwd" + t]), w_inv, sldj, dev(g["ldj_fwd" + t].clone()), reverse=True, **kw)
    assert_close(zr, g["z_rev" + t], what="z_rev")
    ldj_close(lr, g["ldj_rev" + t], "ldj_rev")
    z, ldj = ops.invconv_apply(dev(g["z" + t]), w, sldj, dev(g["ldj_in" + t].clone()))
    assert_close(z, g["z_fwd_plain" + t], what="z plain")
    ldj_close(ldj, g["ldj_fwd_plain" + t], "ldj plain")


def test_logistic_golden():
    from categoricalnf_b200 import ops
    g = load_golden("logistic")
    x = ops.logistic_sample(g.u.shape, "cuda", noise=dev(g.u))
    assert_close(x, g.x, what="sample")
    _, lp = ops.logistic_logprob(dev(g.xs).reshape(1, -1, 1), reduce=False, elementwise=True)
    assert_close(lp.flatten(), g.log_prob, what="log_prob")
    tot, _ = ops.logistic_logprob(dev(g.xs).reshape(1, -1, 1))
    assert_close(tot, g.log_prob.sum().reshape(1), rtol=1e-5, atol=1e-4, what="sum")
    # Philox sampler: moments of Logistic(0, 1/1.81) and reproducibility
    a = ops.logistic_sample((1 << 20,), "cuda", seed=7, offset=3)
    b = ops.logistic_sample((1 << 20,), "cuda", seed=7, offset=3)
    c = ops.logistic_sample((1 << 20,), "cuda", seed=8, offset=3)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1.0021) < 1e-2


@pytest.mark.parametrize("B,S,C,padded", [(64, 256, 16, False), (33, 64, 16, True), (7, 20, 2, True), (5, 38, 6, False), (1, 1, 1, False)])
def test_logistic_logprob_loglik_and_total(B, S, C, padded):
    """``add`` / ``total`` of cnf_logistic_logprob (ABI v3): out = add + log_prob per sample and (sum_b out, B) as a float64
    pair from the same kernel - both the 1024-element row kernel (first two shapes) and the general one - against the oracle."""
    from categoricalnf_b200 import ops
    g = torch.Generator().manual_seed(B + S + C)
    x = torch.randn(B, S, C, generator=g) * 1.5
    ldj = torch.randn(B, generator=g) * 100
    pad = (torch.rand(B, S, generator=g) > 0.3).float() if padded else None
    want = O.logistic_log_prob(x).double()
    if padded:
        want = want * pad.unsqueeze(-1)
    want = want.sum(dim=[1, 2]) + ldj.double()
    total = torch.full((2,), 7.0, dtype=torch.float64, device="cuda")          # stale contents must not survive
    ll, _ = ops.logistic_logprob(dev(x), pad=dev(pad) if padded else None, add=dev(ldj), total=total)
    assert_close(ll, want, rtol=1e-5, atol=2e-4, what="ldj + log_prob")
    assert total[1].item() == float(B)
    assert abs(total[0].item() - want.sum().item()) <= 1e-6 * want.abs().sum().item() + 1e-3
    assert abs(total[0].item() - ll.double().sum().item()) <= 1e-6 * want.abs().sum().item() + 1e-3
    ll2, _ = ops.logistic_logprob(dev(x), pad=dev(pad) if padded else None)     # plain call unchanged
    assert_close(ll2, want - ldj.double(), rtol=1e-5, atol=2e-4, what="log_prob")
    with pytest.raises(ValueError):
        ops.logistic_logprob(dev(x), add=dev(ldj), out=torch.zeros(B, device="cuda"))


@pytest.mark.parametrize("name", ["encode_lm", "encode_mol_nodes", "encode_mol_edges", "encode_virtual"])
def test_categ_encode_decode_golden(name):
    from categoricalnf_b200 import ops
    g = load_golden(name)
    table = O.categ_table(g.embed, g.weight, g.bias)
    pad = dev(g.pad) if g.padded else None
    ldj = dev(g.ldj_in.clone())
    z, ldj, cpl = ops.categ_encode(dev(g.x), dev(table), dev(g.category_prior), ldj, noise=dev(g.u), pad=pad,
                                   beta=g.beta, want_class_prob=True)
    ops.check_status(z.device)
    assert_close(z, g.z, what="z")
    ldj_close(ldj, g.ldj, "ldj")
    assert torch.equal(ops.categ_decode(dev(g.z), dev(table), dev(g.category_prior)).cpu(), g.x_dec)
    assert torch.equal(ops.categ_decode(dev(g.z_rand), dev(table), dev(g.category_prior)).cpu(), g.x_dec_rand)


def test_categ_encode_philox_roundtrip():
    """In-kernel noise: encode -> decode recovers the tokens (posterior mass concentrates on x)."""
    from categoricalnf_b200 import ops
    g = torch.Generator().manual_seed(0)
    V, D, B, S = 51, 16, 64, 256
    table = torch.cat([torch.randn(V, D, generator=g) * 3.0, torch.randn(V, D, generator=g) * 0.3 - 1.0], dim=1)
    prior = torch.log_softmax(torch.zeros(V), 0)
    x = torch.randint(0, V, (B, S), generator=g)
    ldj = torch.zeros(B, device="cuda")
    z, ldj, cpl = ops.categ_encode(dev(x), dev(table), dev(prior), ldj, seed=123, offset=0, want_class_prob=True)
    z2, _, _ = ops.categ_encode(dev(x), dev(table), dev(prior), torch.zeros(B, device="cuda"), seed=123, offset=0)
    assert torch.equal(z, z2)
    dec = ops.categ_decode(z, dev(table), dev(prior))
    assert (dec.cpu() == x).float().mean() > 0.999
    # oracle on the same latent: posterior of the true class, given the noise implied by z
    assert torch.isfinite(ldj).all() and (cpl <= 1e-5).all()


@pytest.mark.parametrize("C,padded", [(16, False), (8, True), (6, True), (32, False)])
def test_invconv_with_actnorm_prologue_and_masked_output(C, padded):
    """cnf_invconv_apply with the ActNorm prologue (+ masked second output) equals ActNorm, 1x1 conv and the
    mask multiply run one after the other."""
    from categoricalnf_b200 import ops
    g = torch.Generator().manual_seed(40 + C)
    B, S = 5, 37
    z = torch.randn(B, S, C, generator=g)
    bias, scales = torch.randn(C, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3
    w = torch.linalg.qr(torch.randn(C, C, generator=g))[0].contiguous()
    sldj = torch.randn(1, generator=g)
    omask = (torch.rand(C, generator=g) > 0.5).float()
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    pad = (torch.arange(S)[None, :] < lens[:, None]).float() if padded else None
    z1, _ = ops.actnorm(dev(z), dev(bias), dev(scales), None, pad=dev(pad))
    z1, _ = ops.invconv_apply(z1, dev(w), dev(sldj), None, pad=dev(pad))
    z2, _, zm = ops.invconv_apply(dev(z), dev(w), dev(sldj), None, pad=dev(pad), pre_actnorm=(dev(bias), dev(scales)),
                                  out_mask=dev(omask))
    assert_close(z2, z1, rtol=1e-5, atol=2e-6, what="actnorm + conv in one pass")
    assert torch.equal(zm, z2 * dev(omask))
    zo, _ = O.actnorm(z, bias.view(1, 1, -1), scales.view(1, 1, -1), pad=pad.unsqueeze(-1) if padded else None)
    zo, _ = O.invconv(zo, w, sldj, pad=pad.unsqueeze(-1) if padded else None)
    assert_close(z2, zo, what="against the oracle")


@pytest.mark.parametrize("B,S", [(16, 256), (9, 131), (64, 100)])
@pytest.mark.parametrize("variant", ["plain", "padded_length", "actnorm_masked", "reverse"])
def test_invconv_tile_kernel_vs_oracle(B, S, variant):
    """C = 16 with >= 1024 positions takes the TMA-tiled kernel (csrc/invconv_tile.cu: 64-byte-swizzled tiles, thread per
    row): full tiles, a ragged last tile (B * S not a multiple of 256), pad mask + lengths, the ActNorm prologue with the
    masked second output, and the reverse direction, each against the oracle."""
    from categoricalnf_b200 import ops
    C = 16
    g = torch.Generator().manual_seed(B * 31 + S)
    z = torch.randn(B, S, C, generator=g)
    w = torch.linalg.qr(torch.randn(C, C, generator=g))[0].contiguous() * 1.1
    sldj = torch.randn(1, generator=g)
    ldj0 = torch.randn(B, generator=g)
    lens = torch.randint(S // 2, S + 1, (B,), generator=g)
    pad = (torch.arange(S)[None, :] < lens[:, None]).float()
    if variant == "plain":
        zo, lo = ops.invconv_apply(dev(z), dev(w), dev(sldj), dev(ldj0.clone()))
        zr, lr = O.invconv(z, w, sldj, ldj0)
    elif variant == "padded_length":
        zo, lo = ops.invconv_apply(dev(z), dev(w), dev(sldj), dev(ldj0.clone()), pad=dev(pad), length=dev(lens.float()))
        zr, lr = O.invconv(z, w, sldj, ldj0, length=lens, pad=pad.unsqueeze(-1))
    elif variant == "reverse":
        w_inv = O.invconv_inverse(w)
        zo, lo = ops.invconv_apply(dev(z), dev(w_inv), dev(sldj), dev(ldj0.clone()), pad=dev(pad), length=dev(lens.float()), reverse=True)
        zr, lr = O.invconv(z, w_inv, sldj, ldj0, reverse=True, length=lens, pad=pad.unsqueeze(-1))
    else:
        bias, scales = torch.randn(C, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3
        omask = torch.tensor([1.0] * 8 + [0.0] * 8)
        zo, lo, zm = ops.invconv_apply(dev(z), dev(w), dev(sldj), dev(ldj0.clone()), pad=dev(pad), pre_actnorm=(dev(bias), dev(scales)),
                                       out_mask=dev(omask))
        za, _ = O.actnorm(z, bias.view(1, 1, -1), scales.view(1, 1, -1), pad=pad.unsqueeze(-1))
        za = za * pad.unsqueeze(-1)
        zr, lr = O.invconv(za, w, sldj, ldj0, pad=pad.unsqueeze(-1))
        assert torch.equal(zm, zo * dev(omask))
    ops.check_status(zo.device)
    assert_close(zo, zr, what="z (%s)" % variant)
    ldj_close(lo, lr, "ldj (%s)" % variant)
