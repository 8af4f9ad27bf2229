"""GPU parity of the fast-path kernels at sizes that select them: the TMA-pipelined mixture
coupling (csrc/mixcdf_pipe.cu) and the thread-per-token categorical encode / decode
(csrc/categ_tpt.cu), each against the CPU oracle on seeded inputs, plus size-independent
properties at BASELINE's full LM size (forward -> inverse round trip, ldj antisymmetry,
linearity of the ldj accumulator in the batch split).

Tolerance: |a-b| <= 1e-4 |b| + 1e-5 on z, 1e-4 relative (+2e-4 abs) on ldj."""
import pytest
import torch

from conftest import assert_close
from oracle import cnf_oracle as O

pytestmark = pytest.mark.gpu


def dev(t):
    return t.cuda() if isinstance(t, torch.Tensor) else t


def _mix_inputs(B, S, C, K, seed, std=0.7):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, S, C, generator=g)
    nn_out = torch.randn(B, S, C * (2 + 3 * K), generator=g) * std
    sf = torch.randn(C, generator=g) * 0.3
    msf = torch.randn(C, K, generator=g) * 0.3
    return z, nn_out, sf, msf


# (B, S, C, K, mask) - all eligible for the pipelined kernel: contiguous transformed run, 16-byte aligned
PIPE_CASES = [
    (6, 37, 16, 8, "half"),      # ragged: 222 positions, tiles straddle samples
    (3, 256, 16, 8, "half"),     # LM layout, S multiple of the tile
    (5, 50, 16, 8, "flip"),      # transformed channels first
    (4, 33, 8, 8, "half"),       # Ct = 4
    (4, 40, 16, 4, "half"),      # K = 4
    (3, 29, 16, 16, "half"),     # K = 16 (two stages)
    (2, 64, 16, 8, "chess"),     # position mask, all 16 channels transformed
]


def _mask(kind, C, S):
    if kind == "half":
        return O.channel_mask(C, 0.5), [1.0] * (C // 2) + [0.0] * (C - C // 2), None
    if kind == "flip":
        m = 1 - O.channel_mask(C, 0.5)
        return m, m.flatten().tolist(), None
    m = O.chess_mask(2)
    return m, None, m.flatten().tolist()


@pytest.mark.parametrize("B,S,C,K,kind", PIPE_CASES)
@pytest.mark.parametrize("padded", [False, True])
def test_mixcdf_pipe_vs_oracle(B, S, C, K, kind, padded):
    from categoricalnf_b200 import ops
    z, nn_out, sf, msf = _mix_inputs(B, S, C, K, seed=B * 1000 + S)
    mask, mc, ms = _mask(kind, C, S)
    pad = None
    if padded:
        length = torch.randint(S // 2, S + 1, (B,), generator=torch.Generator().manual_seed(S))
        pad = (torch.arange(S).view(1, S) < length.view(-1, 1)).float().unsqueeze(-1)
    m = O.expand_mask(mask, z)
    z_ref, ldj_ref, reg_ref = O.mixcdf_coupling(z, nn_out, m, K, sf, msf, pad=pad, reg_max=2.0, reg_factor=0.5, training=True)
    zo, ldj, reg = ops.mixcdf(dev(z), dev(nn_out), K, mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf),
                              mixture_scaling_factor=dev(msf), reg_max=2.0, reg_factor=0.5, training=True, want_reg=True)
    ops.check_status(zo.device)
    assert_close(zo, z_ref, what="z fwd")
    assert_close(ldj, ldj_ref, rtol=1e-4, atol=2e-4, what="ldj fwd")
    assert_close(reg, reg_ref, rtol=1e-4, atol=2e-4, what="reg ldj")
    # inverse of the forward output with the same parameters
    z_inv_ref, ldj_inv_ref, _ = O.mixcdf_coupling(z_ref, nn_out, m, K, sf, msf, pad=pad, reverse=True)
    zi, ldji, _ = ops.mixcdf(dev(z_ref), dev(nn_out), K, mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf),
                             mixture_scaling_factor=dev(msf), reverse=True)
    assert_close(zi, z_inv_ref, what="z inv")
    assert_close(ldji, ldj_inv_ref, rtol=1e-4, atol=2e-4, what="ldj inv")


def test_mixcdf_pipe_matches_generic_kernel(monkeypatch):
    """Both kernels implement the same arithmetic: results agree to a few ulp on the LM layout."""
    import subprocess, sys, os
    code = ("import torch,sys;sys.path.insert(0,%r);from categoricalnf_b200 import ops;"
            "g=torch.Generator().manual_seed(1);z=torch.randn(16,256,16,generator=g).cuda();"
            "nn=(torch.randn(16,256,416,generator=g)*0.7).cuda();"
            "o=ops.mixcdf(z,nn,8,mask_c=[1.]*8+[0.]*8);torch.save([t.cpu() for t in o[:2]],sys.argv[1])")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env in ({}, {"CNF_B200_MIXCDF_GENERIC": "1"}):
        path = "/tmp/cnf_pipe_vs_generic_%d.pt" % len(outs)
        subprocess.run([sys.executable, "-c", code % root, path], check=True, env={**os.environ, **env})
        outs.append(torch.load(path))
    assert_close(outs[0][0], outs[1][0], rtol=2e-6, atol=2e-6, what="z pipe vs generic")
    assert_close(outs[0][1], outs[1][1], rtol=1e-5, atol=1e-4, what="ldj pipe vs generic")


def test_mixcdf_full_size_roundtrip_and_split_linearity():
    """BASELINE size (B=4096, S=256, C=16, K=8): inverse(forward(z)) == z, ldj_inv == -ldj_fwd, and the
    per-sample ldj of the full batch equals that of its two halves run separately."""
    from categoricalnf_b200 import ops
    B, S, C, K = 4096, 256, 16, 8
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(B, S, C, device="cuda", generator=g)
    nn_out = torch.randn(B, S, C * (2 + 3 * K), device="cuda", generator=g) * 0.5
    mc = [1.0] * 8 + [0.0] * 8
    zf, ldj_f, _ = ops.mixcdf(z, nn_out, K, mask_c=mc)
    zr, ldj_r, _ = ops.mixcdf(zf, nn_out, K, mask_c=mc, reverse=True)
    ops.check_status(z.device)
    assert torch.equal(zf[..., :8], z[..., :8])              # conditioner half untouched
    assert_close(zr, z, rtol=1e-4, atol=2e-5, what="round trip")
    assert_close(ldj_r, -ldj_f, rtol=1e-4, atol=1e-2, what="ldj antisymmetry")
    h = B // 2
    _, l0, _ = ops.mixcdf(z[:h], nn_out[:h], K, mask_c=mc)
    _, l1, _ = ops.mixcdf(z[h:], nn_out[h:], K, mask_c=mc)
    assert_close(torch.cat([l0, l1]), ldj_f, rtol=1e-5, atol=1e-3, what="batch split")


@pytest.mark.parametrize("V,D,B,S", [(51, 16, 24, 100), (9, 6, 64, 38), (3, 2, 8, 300), (1, 2, 4, 20), (20, 3, 30, 70),
                                      (51, 4, 16, 140)])
@pytest.mark.parametrize("padded", [False, True])
def test_categ_tpt_vs_oracle(V, D, B, S, padded):
    from categoricalnf_b200 import ops
    g = torch.Generator().manual_seed(V * 100 + D)
    table = torch.cat([torch.randn(V, D, generator=g) * 1.5, torch.randn(V, D, generator=g) * 0.8], dim=1)
    prior = torch.log_softmax(torch.randn(V, generator=g), 0)
    x = torch.randint(0, V, (B, S), generator=g)
    u = torch.rand(B * S, 1, D, generator=g)
    pad = (torch.rand(B, S, 1, generator=g) > 0.2).float() if padded else None
    z_ref, ldj_ref, cpl_ref = O.categ_encode(x, u, table, prior, beta=0.7, pad=pad)
    ldj0 = torch.randn(B, generator=g)
    z, ldj, cpl = ops.categ_encode(dev(x), dev(table), dev(prior), dev(ldj0.clone()), noise=dev(u), pad=dev(pad), beta=0.7,
                                   want_class_prob=True)
    ops.check_status(z.device)
    assert_close(z, z_ref, what="z")
    assert_close(ldj, ldj_ref + ldj0, rtol=1e-4, atol=2e-4, what="ldj")
    assert_close(cpl.reshape(-1), cpl_ref, rtol=1e-4, atol=2e-5, what="class_prob_log")
    # decode: ties aside, the argmax matches the oracle
    dec = ops.categ_decode(dev(z_ref), dev(table), dev(prior)).cpu()
    dec_ref = O.categ_decode(z_ref, table, prior)
    assert (dec == dec_ref).float().mean() > 0.9999


def test_categ_tpt_dominating_other_class():
    """A latent that another class explains e^80 times better than its own: the reference-point
    log-sum-exp must fall back to the exact running-max form (finite, equal to the oracle)."""
    from categoricalnf_b200 import ops
    V, D, B, S = 4, 4, 8, 300
    table = torch.zeros(V, 2 * D)
    table[0, :D] = 60.0        # class 0 sits far away from the others
    table[:, D:] = -2.0        # narrow classes
    prior = torch.log_softmax(torch.zeros(V), 0)
    x = torch.zeros(B, S, dtype=torch.int64)
    x[:, ::2] = 1
    u = torch.rand(B * S, 1, D, generator=torch.Generator().manual_seed(0))
    z_ref, ldj_ref, cpl_ref = O.categ_encode(x, u, table, prior)
    z, ldj, cpl = ops.categ_encode(dev(x), dev(table), dev(prior), torch.zeros(B, device="cuda"), noise=dev(u),
                                   want_class_prob=True)
    assert torch.isfinite(cpl).all()
    assert_close(ldj, ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")


@pytest.mark.parametrize("padded", [False, True])
def test_block_fusion_matches_unfused_and_oracle(padded):
    """ActNorm + 1x1 conv fused into the epilogue of the encode / mixture kernels: module level
    (FlowModel.fuse_blocks) and kernel level (LMDevicePath.forward(fused=...)) against the unfused
    path and the oracle composition."""
    import workload as W
    from categoricalnf_b200 import ops
    B, S = 40, 64            # 2560 tokens: large enough for the fusable encode kernel
    prm = W.data_init_oracle(W.lm_params(seed=5, S=S, blocks=3), seed=5)
    tokens, u = W.lm_tokens(B, S, prm.V, seed=5), W.lm_noise(B, S, prm.D, seed=5)
    dev_ = torch.device("cuda", 0)
    path = W.LMDevicePath(prm, dev_)
    if not padded:
        z_ref, ldj_ref, lp_ref = W.lm_oracle_forward(prm, tokens, u)
        z0, ldj0, lp0 = path.forward(tokens.to(dev_), u_noise=u.to(dev_), fused=False)
        z1, ldj1, lp1 = path.forward(tokens.to(dev_), u_noise=u.to(dev_), fused=True)
        assert ops.categ_encode_fusable(B, S, prm.V, prm.D)
        for z, ldj, lp, what in ((z0, ldj0, lp0, "unfused"), (z1, ldj1, lp1, "fused")):
            assert_close(z, z_ref, what="z " + what)
            assert_close(ldj, ldj_ref, rtol=1e-4, atol=2e-4, what="ldj " + what)
            assert_close(lp, lp_ref, rtol=1e-4, atol=2e-4, what="log prior " + what)
    model, _ = W.build_lm_model(prm, dev_)
    kw = {}
    if padded:
        length = torch.randint(S // 2, S + 1, (B,), generator=torch.Generator().manual_seed(1))
        pad = (torch.arange(S).view(1, S) < length.view(-1, 1)).float().unsqueeze(-1)
        kw = dict(channel_padding_mask=pad.to(dev_), length=length.to(dev_))
    outs = []
    launches = []
    for fuse in (False, True):
        model.fuse_blocks = fuse
        n0 = ops.launch_count()
        with torch.no_grad():
            outs.append(model(tokens.to(dev_), u_noise=u.to(dev_), **kw))
        launches.append(ops.launch_count() - n0)
    assert launches[1] < launches[0], "fusion did not remove launches: %s" % (launches,)
    assert_close(outs[1][0], outs[0][0], rtol=1e-5, atol=2e-6, what="z fused vs unfused (modules)")
    assert_close(outs[1][1], outs[0][1], rtol=1e-5, atol=2e-4, what="ldj fused vs unfused (modules)")
    if not padded:
        assert_close(outs[1][0], z_ref, what="z modules")
        assert_close(outs[1][1], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj modules")


@pytest.mark.parametrize("B,S,C,padded,accumulate", [(5, 64, 16, False, False), (3, 256, 16, True, True), (7, 512, 4, True, False),
                                                     (2, 63, 16, True, False)])
def test_logistic_logprob_row_kernel(B, S, C, padded, accumulate):
    """Per-sample prior log-likelihood (distributions.py:129-137): the warp-per-1024-elements kernel (S*C % 1024 == 0) and,
    last case, the general kernel on the same checks."""
    from categoricalnf_b200 import ops
    g = torch.Generator().manual_seed(B * S)
    x = torch.randn(B, S, C, generator=g) * 3
    x[0, 0, 0], x[0, 0, 1] = 60.0, -60.0
    pad = None
    if padded:
        length = torch.randint(1, S + 1, (B,), generator=g)
        pad = (torch.arange(S)[None, :] < length[:, None]).float()
    ref = O.logistic_log_prob(x.double())
    ref = (ref * pad.unsqueeze(-1).double()).sum(dim=[1, 2]) if padded else ref.sum(dim=[1, 2])
    base = torch.randn(B, generator=g)
    out = base.clone().cuda() if accumulate else None
    res, _ = ops.logistic_logprob(x.cuda(), pad=None if pad is None else pad.cuda(), out=out)
    want = ref + base.double() if accumulate else ref
    assert_close(res, want, rtol=1e-5, atol=1e-3, what="per-sample log-prob")


def test_flow_forward_cuda_graph_replay_matches_eager():
    """GraphedFlowForward: the LM flow (encoding + 3 blocks, stand-in nets) replayed from a CUDA graph gives the eager
    pass' z / ldj / log-likelihood on the same noise, for several batches through one capture, with and without padding
    (two signatures -> two captures), and draws fresh noise when none is given."""
    import workload as W
    from categoricalnf_b200 import ops
    from categoricalnf_b200.layers.flows import GraphedFlowForward
    dev_ = torch.device("cuda", 0)
    prm = W.data_init_oracle(W.lm_params(seed=5, S=64, blocks=3), seed=5)
    model, _ = W.build_lm_model(prm, dev_)
    graphed = GraphedFlowForward(model, log_prior=lambda z, pad: ops.logistic_logprob(z, pad=pad)[0])
    B, S = 16, 64
    with torch.no_grad():
        for trial in range(3):
            tokens, u = W.lm_tokens(B, S, prm.V, seed=40 + trial).to(dev_), W.lm_noise(B, S, prm.D, seed=40 + trial).to(dev_)
            z_e, ldj_e = model(tokens, u_noise=u)
            ll_e = ldj_e + ops.logistic_logprob(z_e)[0]
            z_g, ldj_g, ll_g = graphed(tokens, u_noise=u)
            assert_close(z_g, z_e, rtol=1e-6, atol=1e-6, what="z (trial %d)" % trial)
            assert_close(ldj_g, ldj_e, rtol=1e-6, atol=1e-4, what="ldj (trial %d)" % trial)
            assert_close(ll_g, ll_e, rtol=1e-6, atol=1e-4, what="log-likelihood (trial %d)" % trial)
        assert graphed.captures == 1
        length = torch.randint(S // 2, S + 1, (B,))
        pad = (torch.arange(S)[None, :] < length[:, None]).float().unsqueeze(-1).to(dev_)
        z_e, ldj_e = model(tokens, u_noise=u, channel_padding_mask=pad, length=length.to(dev_))
        z_g, ldj_g, _ = graphed(tokens, u_noise=u, channel_padding_mask=pad, length=length.to(dev_))
        assert_close(z_g, z_e, rtol=1e-6, atol=1e-6, what="z (padded)")
        assert_close(ldj_g, ldj_e, rtol=1e-6, atol=1e-4, what="ldj (padded)")
        assert graphed.captures == 2
        a = graphed(tokens)[1].clone()
        b = graphed(tokens)[1].clone()
        assert torch.isfinite(a).all() and not torch.equal(a, b), "internal noise must differ between replays"
    ops.check_status(dev_, "graphed flow")


# ---- full BASELINE size against the ORACLE on a random subset of samples ---------------------------------------------------
# Samples are independent on the whole path, so the oracle can be run on just the spot-checked samples while the kernel runs
# the full problem (>= 296 CTAs of work for the persistent schedulers, tiles straddling samples, the full grid-stride loops).
def _spot(B, n, seed):
    return torch.randperm(B, generator=torch.Generator().manual_seed(seed))[:n].sort().values


def test_mixcdf_full_size_spot_check_vs_oracle():
    """cnf_mixcdf_fwd / _inv at B 4096 x S 256 x C 16, K 8 (mixcdf_pipe_kernel): 16 random samples vs the oracle,
    forward (z, ldj, regulariser) and inverse; pure-relative ldj deviation asserted next to the usual tolerance."""
    from categoricalnf_b200 import ops
    B, S, C, K = 4096, 256, 16, 8
    g = torch.Generator(device="cuda").manual_seed(11)
    z = torch.randn(B, S, C, device="cuda", generator=g)
    nn_out = torch.randn(B, S, C * (2 + 3 * K), device="cuda", generator=g) * 0.6
    sf = (torch.randn(C, generator=torch.Generator().manual_seed(1)) * 0.3)
    msf = (torch.randn(C, K, generator=torch.Generator().manual_seed(2)) * 0.3)
    mc = [1.0] * 8 + [0.0] * 8
    zf, ldj, reg = ops.mixcdf(z, nn_out, K, mask_c=mc, scaling_factor=dev(sf), mixture_scaling_factor=dev(msf), reg_max=2.0,
                              reg_factor=0.5, training=True, want_reg=True)
    zi, ldji, _ = ops.mixcdf(zf, nn_out, K, mask_c=mc, scaling_factor=dev(sf), mixture_scaling_factor=dev(msf), reverse=True)
    ops.check_status(z.device)
    idx = _spot(B, 16, 3)
    zs, ns = z[idx.cuda()].cpu(), nn_out[idx.cuda()].cpu()
    m = O.expand_mask(O.channel_mask(C, 0.5), zs)
    z_ref, ldj_ref, reg_ref = O.mixcdf_coupling(zs, ns, m, K, sf, msf, reg_max=2.0, reg_factor=0.5, training=True)
    assert_close(zf[idx.cuda()], z_ref, what="z fwd (full size, 16 samples)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj fwd")
    assert_close(reg[idx.cuda()], reg_ref, rtol=1e-4, atol=2e-4, what="reg ldj")
    rel = ((ldj[idx.cuda()].cpu().double() - ldj_ref.double()).abs() / ldj_ref.double().abs()).max().item()
    assert rel <= 1e-4, "pure relative ldj deviation %.3e" % rel
    z_inv_ref, ldj_inv_ref, _ = O.mixcdf_coupling(zf[idx.cuda()].cpu(), ns, m, K, sf, msf, reverse=True)
    assert_close(zi[idx.cuda()], z_inv_ref, what="z inv (full size, 16 samples)")
    assert_close(ldji[idx.cuda()], ldj_inv_ref, rtol=1e-4, atol=2e-4, what="ldj inv")


def test_linear_mixcdf_full_size_spot_check_vs_oracle():
    """cnf_linear_mixcdf_fwd (projection + transform in one tcgen05 kernel) at B 4096 x S 256 x C 16, K 8, H 16."""
    from categoricalnf_b200 import ops
    B, S, C, K, H = 4096, 256, 16, 8, 16
    g = torch.Generator(device="cuda").manual_seed(12)
    z = torch.randn(B, S, C, device="cuda", generator=g)
    feat = z * torch.tensor([1.0] * 8 + [0.0] * 8, device="cuda")
    gw = torch.Generator().manual_seed(4)
    w = torch.randn(C * (2 + 3 * K), H, generator=gw) * (0.5 / (H / 2) ** 0.5)
    b = torch.randn(C * (2 + 3 * K), generator=gw) * 0.1
    sf, msf = torch.randn(C, generator=gw) * 0.3, torch.randn(C, K, generator=gw) * 0.3
    mc = [1.0] * 8 + [0.0] * 8
    if not ops.linear_mixcdf_fusable(z, feat, dev(w), K, mask_c=mc):
        pytest.skip("shape not fusable")
    zf, ldj, _ = ops.linear_mixcdf(z, feat, dev(w), dev(b), K, mask_c=mc, scaling_factor=dev(sf), mixture_scaling_factor=dev(msf))
    ops.check_status(z.device)
    idx = _spot(B, 16, 5)
    zs = z[idx.cuda()].cpu()
    m = O.expand_mask(O.channel_mask(C, 0.5), zs)
    nn_out = torch.nn.functional.linear((zs * m).double(), w.double(), b.double()).float()
    z_ref, ldj_ref, _ = O.mixcdf_coupling(zs, nn_out, m, K, sf, msf)
    assert_close(zf[idx.cuda()], z_ref, what="z (full size, 16 samples)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")
    rel = ((ldj[idx.cuda()].cpu().double() - ldj_ref.double()).abs() / ldj_ref.double().abs()).max().item()
    assert rel <= 1e-4, "pure relative ldj deviation %.3e" % rel


def test_categ_encode_full_size_spot_check_vs_oracle():
    """cnf_categ_encode at tokens [4096, 256], V 51, d 16 (categ_encode_tpt_kernel<16>) on explicit noise."""
    from categoricalnf_b200 import ops
    import workload as W
    B, S, V, D = 4096, 256, 51, 16
    prm = W.lm_params(seed=0)
    tokens = W.lm_tokens(B, S, V, seed=3)
    u = torch.rand(B * S, 1, D, generator=torch.Generator().manual_seed(9))
    ldj0 = torch.zeros(B, device="cuda")
    z, ldj, _ = ops.categ_encode(dev(tokens), dev(prm.table()), dev(prm.prior), ldj0, noise=dev(u))
    ops.check_status(z.device)
    idx = _spot(B, 16, 7)
    u_s = u.view(B, S, 1, D)[idx].reshape(-1, 1, D)
    z_ref, ldj_ref, _ = O.categ_encode(tokens[idx], u_s, prm.table(), prm.prior)
    assert_close(z[idx.cuda()], z_ref, what="z (full size, 16 samples)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")


def test_lm_flow_full_size_spot_check_vs_oracle():
    """The whole LM path of bench.py at its full size (tokens [4096,256] -> encode -> 8 x [ActNorm, 1x1 conv, mixture coupling]
    -> prior) through the kernel-level path with fused blocks: 8 random samples against the oracle's composition, incl.
    bits/dim within 1e-3."""
    import workload as W
    B = 4096
    prm = W.data_init_oracle(W.lm_params(seed=0), seed=0)
    path = W.LMDevicePath(prm, torch.device("cuda", 0))
    tokens = W.lm_tokens(B, prm.S, prm.V, seed=21)
    u = torch.rand(B * prm.S, 1, prm.D, generator=torch.Generator().manual_seed(22))
    z, ldj, lp = path.forward(tokens.cuda(), u_noise=u.cuda())
    idx = _spot(B, 8, 23)
    u_s = u.view(B, prm.S, 1, prm.D)[idx].reshape(-1, 1, prm.D)
    z_ref, ldj_ref, lp_ref = W.lm_oracle_forward(prm, tokens[idx], u_s)
    assert_close(z[idx.cuda()], z_ref, what="z (full size, 8 samples)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj")
    assert_close(lp[idx.cuda()], lp_ref, rtol=1e-4, atol=2e-4, what="log prior")
    bpd, bpd_ref = W.bits_per_dim(ldj[idx.cuda()].cpu(), lp[idx.cuda()].cpu(), prm.S), W.bits_per_dim(ldj_ref, lp_ref, prm.S)
    assert abs(bpd - bpd_ref) <= 1e-3


# ---- lane-group TMA pipeline (csrc/mixcdf_gpipe.cu): any K, unaligned transformed runs -----------------------------------
# (B, S, C, K, mask): K = 64 is the reference's LM default; C = 6 / K = 16 and C = 2 / K = 8 are GraphCNF's node and edge
# flows (runs of 600 and 104 bytes at offsets that are not multiples of 16); odd B * S leaves a ragged last tile whose z rows
# are not a multiple of 16 bytes.
GPIPE_CASES = [
    (3, 50, 16, 64, "half"),
    (2, 40, 16, 32, "half"),
    (5, 38, 6, 16, "half"),
    (3, 33, 6, 16, "flip"),
    (7, 21, 2, 8, "half"),
    (4, 20, 2, 8, "chess"),
    (3, 19, 4, 10, "half"),
    (2, 300, 6, 16, "half"),
    (3, 17, 8, 3, "half"),
]


def _mask_any(kind, C):
    if kind == "chess":
        m = O.chess_mask(2)
        return m, None, m.flatten().tolist()
    m = O.channel_mask(C, 0.5)
    if kind == "flip":
        m = 1 - m
    return m, m.flatten().tolist(), None


@pytest.mark.parametrize("B,S,C,K,kind", GPIPE_CASES)
@pytest.mark.parametrize("padded", [False, True])
def test_mixcdf_gpipe_vs_oracle(B, S, C, K, kind, padded):
    from categoricalnf_b200 import ops
    z, nn_out, sf, msf = _mix_inputs(B, S, C, K, seed=B * 977 + S + K)
    mask, mc, ms = _mask_any(kind, C)
    assert ops.mixcdf_path(dev(z), dev(nn_out), K, mask_c=mc, mask_s=ms) == "gpipe"
    pad = None
    if padded:
        length = torch.randint(S // 2, S + 1, (B,), generator=torch.Generator().manual_seed(S))
        pad = (torch.arange(S).view(1, S) < length.view(-1, 1)).float().unsqueeze(-1)
    m = O.expand_mask(mask, z)
    z_ref, ldj_ref, reg_ref = O.mixcdf_coupling(z, nn_out, m, K, sf, msf, pad=pad, reg_max=2.0, reg_factor=0.5, training=True)
    zo, ldj, reg = ops.mixcdf(dev(z), dev(nn_out), K, mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf),
                              mixture_scaling_factor=dev(msf), reg_max=2.0, reg_factor=0.5, training=True, want_reg=True)
    ops.check_status(zo.device)
    assert_close(zo, z_ref, what="z fwd")
    assert_close(ldj, ldj_ref, rtol=1e-4, atol=2e-4, what="ldj fwd")
    assert_close(reg, reg_ref, rtol=1e-4, atol=2e-4, what="reg ldj")
    z_inv_ref, ldj_inv_ref, _ = O.mixcdf_coupling(z_ref, nn_out, m, K, sf, msf, pad=pad, reverse=True)
    ldj0 = torch.randn(B, generator=torch.Generator().manual_seed(1))
    zi, ldji, _ = ops.mixcdf(dev(z_ref), dev(nn_out), K, mask_c=mc, mask_s=ms, pad=dev(pad), scaling_factor=dev(sf),
                             mixture_scaling_factor=dev(msf), reverse=True, ldj=dev(ldj0.clone()))      # accumulate mode
    assert_close(zi, z_inv_ref, what="z inv")
    assert_close(ldji, ldj_inv_ref + ldj0, rtol=1e-4, atol=2e-4, what="ldj inv (accumulated)")


def test_mixcdf_paths():
    """Dispatch: compile-time pipeline for the LM layout, lane-group pipeline for other K / unaligned runs, staged generic kernel
    for non-contiguous masks and rows that are not a multiple of 16 bytes."""
    from categoricalnf_b200 import ops
    def path(C, K, mc=None, ms=None, B=2, S=8):
        z = torch.zeros(B, S, C, device="cuda")
        return ops.mixcdf_path(z, torch.zeros(B, S, C * (2 + 3 * K), device="cuda"), K, mask_c=mc, mask_s=ms)
    half = lambda C: [1.0] * (C // 2) + [0.0] * (C - C // 2)
    assert path(16, 8, half(16)) == "pipe"
    assert path(16, 64, half(16)) == "gpipe"
    assert path(6, 16, half(6)) == "gpipe" and path(2, 8, half(2)) == "gpipe"
    assert path(4, 8, [1.0, 0.0, 1.0, 0.0]) == "generic"          # transformed channels not contiguous
    assert path(3, 3, [1.0, 0.0, 0.0]) == "generic"               # row of 33 floats: not a multiple of 16 bytes


def test_mixcdf_gpipe_k64_large_batch_split_and_roundtrip():
    """K = 64 at B 256 x S 256 x C 16 (>= 2 CTAs of tiles per SM on the persistent grid): round trip, ldj antisymmetry, batch
    split, and 8 random samples against the oracle."""
    from categoricalnf_b200 import ops
    B, S, C, K = 256, 256, 16, 64
    g = torch.Generator(device="cuda").manual_seed(3)
    z = torch.randn(B, S, C, device="cuda", generator=g)
    nn_out = torch.randn(B, S, C * (2 + 3 * K), device="cuda", generator=g) * 0.6
    mc = [1.0] * 8 + [0.0] * 8
    assert ops.mixcdf_path(z, nn_out, K, mask_c=mc) == "gpipe"
    zf, ldj, _ = ops.mixcdf(z, nn_out, K, mask_c=mc)
    zr, ldjr, _ = ops.mixcdf(zf, nn_out, K, mask_c=mc, reverse=True)
    ops.check_status(z.device)
    assert torch.equal(zf[..., :8], z[..., :8])
    assert_close(zr, z, rtol=1e-4, atol=2e-5, what="round trip")
    assert_close(ldjr, -ldj, rtol=1e-4, atol=1e-2, what="ldj antisymmetry")
    h = B // 2
    _, l0, _ = ops.mixcdf(z[:h], nn_out[:h], K, mask_c=mc)
    assert_close(l0, ldj[:h], rtol=1e-5, atol=1e-3, what="batch split")
    idx = _spot(B, 8, 4)
    zs, ns = z[idx.cuda()].cpu(), nn_out[idx.cuda()].cpu()
    m = O.expand_mask(O.channel_mask(C, 0.5), zs)
    zero = torch.zeros(C), torch.zeros(C, K)       # scaling factors omitted = 0, i.e. bounds e^0 = 1 (a freshly built layer)
    z_ref, ldj_ref, _ = O.mixcdf_coupling(zs, ns, m, K, zero[0], zero[1])
    assert_close(zf[idx.cuda()], z_ref, what="z fwd (8 samples)")
    assert_close(ldj[idx.cuda()], ldj_ref, rtol=1e-4, atol=2e-4, what="ldj fwd")
