"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Usage (authoring container only - the GPU box has no /root/reference):

    python tests/golden/make_golden.py [/root/reference] [--only=node_edge]

For every hot-path function of SURVEY.md section 8a it instantiates the reference
module from the reference checkout, feeds it seeded synthetic inputs on the CPU, and
stores inputs + parameters + outputs as a small ``.npz``.  The oracle
(``oracle/cnf_oracle.py``) and the CUDA path are both checked against these files.
The only shim is an empty ``matplotlib`` module (absent here, imported but unused by
``layers/flows/distributions.py:9``).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

_ARGS = [a for a in sys.argv[1:] if not a.startswith("--only")]
ONLY = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]      # --only=node_edge : just that group
REF = _ARGS[0] if _ARGS else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_matplotlib():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.colors"].hsv_to_rgb = lambda x: x
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]


_stub_matplotlib()
sys.path.insert(0, REF)

from layers.flows.coupling_layer import CouplingLayer                                # noqa: E402
from layers.flows.mixture_cdf_layer import MixtureCDFCoupling                        # noqa: E402
from layers.flows.activation_normalization import ActNormFlow, ExtActNormFlow        # noqa: E402
from layers.flows.permutation_layers import InvertibleConv                           # noqa: E402
from layers.flows.distributions import LogisticDistribution                          # noqa: E402
from layers.flows.autoregressive_coupling import AutoregressiveMixtureCDFCoupling    # noqa: E402
from layers.flows.flow_model import FlowModel                                        # noqa: E402
from layers.categorical_encoding.linear_encoding import LinearCategoricalEncoding    # noqa: E402
from layers.networks.help_layers import SimpleLinearLayer                            # noqa: E402


def npy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: npy(v) for k, v in arrays.items()})
    print("wrote %-34s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024))


class _Recorder(nn.Module):
    """Black-box coupling net that returns a preset tensor (so nn_out is an explicit input)."""

    def __init__(self, out):
        super().__init__()
        self.out = out

    def forward(self, x, **kwargs):
        return self.out


def lengths_to_pad(length, S):
    return (torch.arange(S).view(1, S) < length.view(-1, 1)).float().unsqueeze(-1)


# ---------------------------------------------------------------------------
def gold_mixcdf(name, B, S, C, K, *, seed, nn_std=0.5, sf_std=0.0, chess=False, flip=False,
                padded=False, reg_max=-1.0, reg_factor=1.0, training=True, ratio=0.5, z_std=1.0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, S, C, generator=g) * z_std
    nn_out = torch.randn(B, S, C * (2 + 3 * K), generator=g) * nn_std
    mask = CouplingLayer.create_chess_mask() if chess else CouplingLayer.create_channel_mask(C, ratio=ratio)
    if flip:
        mask = 1 - mask
    layer = MixtureCDFCoupling(c_in=C, mask=mask, model_func=lambda c_out: _Recorder(nn_out),
                               num_mixtures=K, regularizer_max=reg_max, regularizer_factor=reg_factor)
    layer.scaling_factor.data = torch.randn(C, generator=g) * sf_std
    layer.mixture_scaling_factor.data = torch.randn(C, K, generator=g) * sf_std
    layer.train(training)
    kw = {}
    length = torch.full((B,), S, dtype=torch.long)
    if padded:
        length = torch.randint(max(1, S // 3), S + 1, (B,), generator=g)
        length[0] = S
        kw["channel_padding_mask"] = lengths_to_pad(length, S)
        z = z * kw["channel_padding_mask"]
    with torch.no_grad():
        z_fwd, ldj_fwd, det = layer(z, reverse=False, **kw)
        z_rev, ldj_rev, _ = layer(z_fwd, reverse=True, **kw)
        # reverse pass on an independent latent as well (sampling direction)
        z_lat = torch.randn(B, S, C, generator=g) * 1.5
        if padded:
            z_lat = z_lat * kw["channel_padding_mask"]
        z_smp, ldj_smp, _ = layer(z_lat, reverse=True, **kw)
    save(name, z=z, nn_out=nn_out, mask=mask, K=K, sf=layer.scaling_factor.data,
         msf=layer.mixture_scaling_factor.data, length=length, padded=int(padded),
         reg_max=reg_max, reg_factor=reg_factor, training=int(training),
         z_fwd=z_fwd, ldj_fwd=ldj_fwd, reg_ldj=det["regularizer_ldj"],
         z_rev=z_rev, ldj_rev=ldj_rev, z_lat=z_lat, z_smp=z_smp, ldj_smp=ldj_smp)


def gold_mixcdf_tails(name, seed, right_tail=False):
    """Extreme inputs: far tails of every component, sharp and wide components (exercises the
    clamps at mixture_cdf_layer.py:197-198).

    Where the CDF exceeds 1 - 1e-13 the reference's own output is round-off noise of the float64
    `1 - exp(log_cdf)` (its log(1-F) jumps between -36.7 and the -50.66 clamp from one rounding
    to the next), so there is nothing to pin.  The main fixture therefore moves such elements
    left until 1 - CDF >= 1e-11; `right_tail=True` keeps them for a sanity-only test."""
    from layers.flows.mixture_cdf_layer import mixture_log_cdf
    g = torch.Generator().manual_seed(seed)
    B, S, C, K = 2, 24, 4, 4
    z = torch.randn(B, S, C, generator=g) * 12.0
    z[0, 0:6, 2] = torch.tensor([-60.0, -140.0, -300.0, -45.0, -90.0, -35.0])   # transformed channels
    z[0, 6:12, 3] = torch.tensor([-25.0, -50.0, -75.0, -100.0, -1000.0, -20.0])
    z[1, 0:6, 2] = torch.tensor([20.0, 35.0, 60.0, 140.0, 25.0, 30.0])
    z[1, 6:10, 3] = torch.tensor([15.0, 45.0, 90.0, 300.0])
    nn_out = torch.randn(B, S, C * (2 + 3 * K), generator=g) * 2.0
    mask = CouplingLayer.create_channel_mask(C)
    layer = MixtureCDFCoupling(c_in=C, mask=mask, model_func=lambda c_out: _Recorder(nn_out), num_mixtures=K)
    layer.scaling_factor.data = torch.tensor([0.5, -0.5, 1.0, 0.0])
    layer.mixture_scaling_factor.data = torch.randn(C, K, generator=g) * 0.8
    layer.eval()
    with torch.no_grad():
        m3 = mask.unsqueeze(0)
        prm = MixtureCDFCoupling.get_mixt_params(nn_out, m3, K, layer.scaling_factor, layer.mixture_scaling_factor)
        for _ in range(200):
            surv = 1.0 - mixture_log_cdf(z.double(), prm[2], prm[3], prm[4]).exp()
            bad = (surv < 1e-11) & (m3 == 0)
            if right_tail or not bad.any():
                break
            z = torch.where(bad, z - torch.clamp(0.1 * z.abs(), min=0.5), z)
        z_fwd, ldj_fwd, det = layer(z, reverse=False)
    save(name, z=z, nn_out=nn_out, mask=mask, K=K, sf=layer.scaling_factor.data,
         msf=layer.mixture_scaling_factor.data, length=torch.full((B,), S), padded=0,
         reg_max=-1.0, reg_factor=1.0, training=0, z_fwd=z_fwd, ldj_fwd=ldj_fwd,
         reg_ldj=det["regularizer_ldj"])


def gold_mixcdf_inv_underflow():
    """Inverse pass whose first bisection midpoint lands where 1-F and f underflow in float32 (found on
    the GPU: a Newton step computed from a flushed density must not be accepted).  Parameters come from
    a per-position Linear like tools/debug_nan.py; K = 16, C = 8."""
    B, S, C, K, H = 3, 77, 8, 16, 64
    g = torch.Generator().manual_seed(B * 131 + S * 7 + C + K + H)
    PN = 2 + 3 * K
    z_lat = torch.randn(B, S, C, generator=g) * 1.2
    feats = torch.randn(B, S, H, generator=g)
    w = torch.randn(C * PN, H, generator=g) * (0.5 / H ** 0.5)
    b = torch.randn(C * PN, generator=g) * 0.1
    sf = torch.randn(C, generator=g) * 0.3
    msf = torch.randn(C, K, generator=g) * 0.3
    nn_out = (feats.double() @ w.double().t() + b.double()).float()
    mask = CouplingLayer.create_channel_mask(C)
    layer = MixtureCDFCoupling(c_in=C, mask=mask, model_func=lambda c_out: _Recorder(nn_out), num_mixtures=K)
    layer.scaling_factor.data = sf
    layer.mixture_scaling_factor.data = msf
    layer.eval()
    with torch.no_grad():
        z_smp, ldj_smp, _ = layer(z_lat, reverse=True)
    save("mixcdf_inv_underflow", nn_out=nn_out, mask=mask, K=K, sf=sf, msf=msf, z_lat=z_lat, z_smp=z_smp, ldj_smp=ldj_smp)


def gold_mixcdf_selftest():
    """The reference's own __main__ smoke block (mixture_cdf_layer.py:279-302)."""
    torch.manual_seed(42)
    B, S, C, K, H = 8, 16, 4, 10, 128
    captured = {}

    def model_func(c_out):
        net = nn.Sequential(nn.Linear(C, H), nn.ReLU(), nn.Linear(H, c_out))
        net.register_forward_hook(lambda m, i, o: captured.__setitem__("nn_out", o.detach().clone()))
        return net

    mask = CouplingLayer.create_channel_mask(C)
    layer = MixtureCDFCoupling(c_in=C, mask=mask, model_func=model_func, block_type="Linear net", num_mixtures=K)
    x = torch.randn(size=(B, S, C))
    with torch.no_grad():
        z_fwd, ldj_fwd, _ = layer(z=x, reverse=False)
        nn_fwd = captured["nn_out"]
        z_rev, ldj_rev, _ = layer(z=z_fwd, reverse=True)
        nn_rev = captured["nn_out"]
    print("   reference self-test: max reconstruction err %.3e, max ldj err %.3e"
          % ((x - z_rev).abs().max(), (ldj_fwd + ldj_rev).abs().max()))
    save("mixcdf_selftest", z=x, nn_out=nn_fwd, nn_out_rev=nn_rev, mask=mask, K=K,
         sf=layer.scaling_factor.data, msf=layer.mixture_scaling_factor.data,
         z_fwd=z_fwd, ldj_fwd=ldj_fwd, z_rev=z_rev, ldj_rev=ldj_rev)


def gold_autoregressive(seed):
    g = torch.Generator().manual_seed(seed)
    B, S, C, K = 3, 20, 3, 51
    z = torch.randn(B, S, C, generator=g)
    nn_out = torch.randn(B, S, C * (2 + 3 * K), generator=g) * 0.7
    layer = AutoregressiveMixtureCDFCoupling(c_in=C, model_func=lambda c_out: _Recorder(nn_out), num_mixtures=K)
    layer.scaling_factor.data = torch.randn(C, generator=g) * 0.2
    layer.mixture_scaling_factor.data = torch.randn(C, K, generator=g) * 0.2
    ldj0 = torch.randn(B, generator=g)
    length = torch.tensor([20, 13, 7])
    pad = lengths_to_pad(length, S)
    with torch.no_grad():
        z_out, ldj = layer(z, ldj=ldj0.clone(), channel_padding_mask=pad)
    save("autoregressive_mixcdf", z=z, nn_out=nn_out, K=K, sf=layer.scaling_factor.data,
         msf=layer.mixture_scaling_factor.data, ldj_in=ldj0, pad=pad, z_out=z_out, ldj_out=ldj)


def gold_affine(seed):
    g = torch.Generator().manual_seed(seed)
    B, S, C = 5, 9, 6
    z = torch.randn(B, S, C, generator=g)
    nn_out = torch.randn(B, S, 2 * C, generator=g)
    mask = CouplingLayer.create_channel_mask(C)
    layer = CouplingLayer(c_in=C, mask=mask, model_func=lambda c_out: _Recorder(nn_out))
    layer.scaling_factor.data = torch.randn(C, generator=g) * 0.5
    ldj0 = torch.randn(B, generator=g)
    with torch.no_grad():
        z_fwd, ldj_fwd = layer(z, ldj=ldj0.clone())
        z_rev, ldj_rev = layer(z_fwd, ldj=ldj_fwd.clone(), reverse=True)
    save("affine_coupling", z=z, nn_out=nn_out, mask=mask, sf=layer.scaling_factor.data, ldj_in=ldj0,
         z_fwd=z_fwd, ldj_fwd=ldj_fwd, z_rev=z_rev, ldj_rev=ldj_rev)
    # token-wise layout used inside the linear-flow encoding: [B*S, 1, D]
    z = torch.randn(40, 1, 4, generator=g)
    nn_out = torch.randn(40, 1, 8, generator=g)
    mask = CouplingLayer.create_channel_mask(4)
    layer = CouplingLayer(c_in=4, mask=mask, model_func=lambda c_out: _Recorder(nn_out))
    with torch.no_grad():
        z_fwd, ldj_fwd = layer(z)
        z_rev, ldj_rev = layer(z_fwd, ldj=ldj_fwd.clone(), reverse=True)
    save("affine_coupling_tokens", z=z, nn_out=nn_out, mask=mask, sf=layer.scaling_factor.data,
         ldj_in=torch.zeros(40), z_fwd=z_fwd, ldj_fwd=ldj_fwd, z_rev=z_rev, ldj_rev=ldj_rev)


def gold_actnorm(seed):
    g = torch.Generator().manual_seed(seed)
    B, S, C = 6, 11, 5
    z = torch.randn(B, S, C, generator=g) * 2 + 0.7
    length = torch.tensor([11, 3, 7, 11, 1, 9])
    pad = lengths_to_pad(length, S)
    layer = ActNormFlow(C)
    layer.bias.data = torch.randn(1, 1, C, generator=g)
    layer.scales.data = torch.randn(1, 1, C, generator=g) * 0.4
    out = {}
    with torch.no_grad():
        for tag, kw in (("plain", {}), ("len", {"length": length, "channel_padding_mask": pad}),
                        ("padonly", {"channel_padding_mask": pad})):
            ldj0 = torch.randn(B, generator=g)
            zf, lf = layer(z, ldj=ldj0.clone(), **kw)
            zr, lr = layer(zf, ldj=lf.clone(), reverse=True, **kw)
            out.update({"ldj_in_" + tag: ldj0, "z_fwd_" + tag: zf, "ldj_fwd_" + tag: lf,
                        "z_rev_" + tag: zr, "ldj_rev_" + tag: lr})
        init = ActNormFlow(C)
        _quiet(init.data_init_forward, z, channel_padding_mask=pad)
        init2 = ActNormFlow(C)
        _quiet(init2.data_init_forward, z)
    save("actnorm", z=z, length=length, pad=pad, bias=layer.bias.data, scales=layer.scales.data,
         init_bias_pad=init.bias.data, init_scales_pad=init.scales.data,
         init_bias=init2.bias.data, init_scales=init2.scales.data, **out)


def _quiet(fn, *a, **kw):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def gold_ext_actnorm(seed):
    g = torch.Generator().manual_seed(seed)
    N, D, E = 30, 4, 16
    z = torch.randn(N, 1, D, generator=g)
    ext = torch.randn(N, 1, E, generator=g)
    net = SimpleLinearLayer(c_in=E, c_out=2 * D, data_init=True)
    net.layer.weight.data = torch.randn(2 * D, E, generator=g) * 0.3
    net.layer.bias.data = torch.randn(2 * D, generator=g) * 0.3
    layer = ExtActNormFlow(c_in=D, net=net)
    pad = (torch.rand(N, 1, 1, generator=g) > 0.3).float()
    ldj0 = torch.randn(N, generator=g)
    with torch.no_grad():
        zf, lf = layer(z, ldj0.clone(), ext_input=ext, channel_padding_mask=pad)
        zr, lr = layer(zf, lf.clone(), reverse=True, ext_input=ext, channel_padding_mask=pad)
        zf2, lf2 = layer(z, ldj0.clone(), ext_input=ext)
    save("ext_actnorm", z=z, ext=ext, weight=net.layer.weight.data, bias=net.layer.bias.data, pad=pad,
         ldj_in=ldj0, z_fwd=zf, ldj_fwd=lf, z_rev=zr, ldj_rev=lr, z_fwd_nopad=zf2, ldj_fwd_nopad=lf2)


def gold_invconv(seed):
    out = {}
    for C in (2, 6, 16):
        np.random.seed(seed + C)
        g = torch.Generator().manual_seed(seed + C)
        layer = InvertibleConv(c_in=C)
        # perturb away from the orthogonal init so that log_s != 0
        layer.l.data += torch.randn(C, C, generator=g) * 0.1
        layer.u.data += torch.randn(C, C, generator=g) * 0.1
        layer.log_s.data += torch.randn(C, generator=g) * 0.2
        B, S = 4, 7
        z = torch.randn(B, S, C, generator=g)
        length = torch.tensor([7, 2, 5, 7])
        pad = lengths_to_pad(length, S)
        ldj0 = torch.randn(B, generator=g)
        layer.train()
        with torch.no_grad():
            w, sldj = layer._get_weight("cpu", inverse=False)
            w_inv, _ = layer._get_weight("cpu", inverse=True)
            zf, lf = layer(z, ldj=ldj0.clone(), length=length, channel_padding_mask=pad)
            zr, lr = layer(zf, ldj=lf.clone(), reverse=True, length=length, channel_padding_mask=pad)
            zf2, lf2 = layer(z, ldj=ldj0.clone())
        t = "_c%d" % C
        out.update({"p" + t: layer.p, "sign_s" + t: layer.sign_s, "l" + t: layer.l.data, "u" + t: layer.u.data,
                    "log_s" + t: layer.log_s.data, "w" + t: w, "w_inv" + t: w_inv, "sldj" + t: sldj,
                    "z" + t: z, "length" + t: length, "pad" + t: pad, "ldj_in" + t: ldj0,
                    "z_fwd" + t: zf, "ldj_fwd" + t: lf, "z_rev" + t: zr, "ldj_rev" + t: lr,
                    "z_fwd_plain" + t: zf2, "ldj_fwd_plain" + t: lf2})
    save("invconv", **out)


def gold_logistic(seed):
    g = torch.Generator().manual_seed(seed)
    dist = LogisticDistribution(mu=0.0, sigma=1.0)
    u = torch.rand(64, 1, 5, generator=g)
    u[0, 0, 0], u[1, 0, 0] = 0.0, 1.0 - 2 ** -24
    rec = {}
    orig = dist.distribution.sample
    dist.distribution.sample = lambda sample_shape=torch.Size(): u
    x = dist.sample(shape=(64, 1, 5))
    dist.distribution.sample = orig
    xs = torch.cat([x.flatten(), torch.tensor([0.0, 30.0, -30.0, 1e-3, 8.0])])
    save("logistic", u=u, x=x, xs=xs, log_prob=dist.log_prob(xs), **rec)


def gold_encoding(name, B, S, V, D, *, seed, padded=False, beta=1.0, prior_std=0.0, training=False):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    prior = torch.randn(V, generator=g) * prior_std if prior_std > 0 else None
    enc = LinearCategoricalEncoding(num_dimensions=D, flow_config={"num_flows": 0}, vocab_size=V,
                                    category_prior=prior)
    lin = enc.flow_layers[0].pred_net.layer
    lin.weight.data = torch.randn(2 * D, 64, generator=g) * 0.25     # non-zero scale rows
    lin.bias.data = torch.randn(2 * D, generator=g) * 0.3
    enc.train(training)
    x = torch.randint(0, V, (B, S), generator=g)
    kw = {}
    pad = torch.ones(B, S, 1)
    if padded:
        length = torch.randint(1, S + 1, (B,), generator=g)
        pad = lengths_to_pad(length, S)
        kw["channel_padding_mask"] = pad
    noise = {}
    orig = enc.prior_distribution.distribution.sample

    def rec_sample(sample_shape=torch.Size()):
        noise["u"] = torch.rand(sample_shape, generator=g)
        return noise["u"]

    enc.prior_distribution.distribution.sample = rec_sample
    ldj0 = torch.randn(B, generator=g)
    with torch.no_grad():
        z, ldj, _ = enc(x, ldj=ldj0.clone(), beta=beta, **kw)
        enc.prior_distribution.distribution.sample = orig
        x_dec, _, _ = enc(z, reverse=True, **kw)
        z_rand = torch.randn(B, S, D, generator=g)
        x_dec_rand, _, _ = enc(z_rand, reverse=True)
    save(name, x=x, u=noise["u"], V=V, D=D, beta=beta, pad=pad, padded=int(padded), ldj_in=ldj0,
         embed=enc.embed_layer.weight.data, weight=lin.weight.data, bias=lin.bias.data,
         category_prior=enc.category_prior, z=z, ldj=ldj, x_dec=x_dec, z_rand=z_rand, x_dec_rand=x_dec_rand)


def gold_lm_flow(seed):
    """Small version of BASELINE config 2: encode -> 3 x [ActNorm, InvConv, MixtureCDF] -> prior,
    run through the reference's FlowModel container (flow_model.py:25-53)."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    B, S, V, D, K, NB, H = 4, 24, 51, 16, 8, 3, 32
    enc = LinearCategoricalEncoding(num_dimensions=D, flow_config={"num_flows": 0}, vocab_size=V)
    lin = enc.flow_layers[0].pred_net.layer
    lin.weight.data = torch.randn(2 * D, 64, generator=g) * 0.2
    lin.bias.data = torch.randn(2 * D, generator=g) * 0.3
    captured = []

    def model_func(c_out):
        net = nn.Sequential(nn.Linear(D, H), nn.GELU(), nn.Linear(H, c_out))
        net.register_forward_hook(lambda m, i, o: captured.append(o.detach().clone()))
        return net

    layers = [enc]
    mask = CouplingLayer.create_channel_mask(D)
    for i in range(NB):
        an, ic = ActNormFlow(D), InvertibleConv(D)
        an.bias.data = torch.randn(1, 1, D, generator=g) * 0.2
        an.scales.data = torch.randn(1, 1, D, generator=g) * 0.2
        ic.log_s.data += torch.randn(D, generator=g) * 0.1
        cp = MixtureCDFCoupling(c_in=D, mask=mask if i % 2 == 0 else 1 - mask, model_func=model_func, num_mixtures=K)
        cp.scaling_factor.data = torch.randn(D, generator=g) * 0.2
        cp.mixture_scaling_factor.data = torch.randn(D, K, generator=g) * 0.2
        layers += [an, ic, cp]
    model = _quiet(FlowModel, layers)
    model.eval()
    x = torch.randint(0, V, (B, S), generator=g)
    length = torch.tensor([24, 17, 9, 24])
    pad = lengths_to_pad(length, S)
    noise = {}

    def rec_sample(sample_shape=torch.Size()):
        noise["u"] = torch.rand(sample_shape, generator=g)
        return noise["u"]

    enc.prior_distribution.distribution.sample = rec_sample
    with torch.no_grad():
        z, ldj = model(x, reverse=False, length=length, channel_padding_mask=pad)
        logp = (LogisticDistribution().log_prob(z) * pad).sum(dim=[1, 2])
    arrays = dict(x=x, u=noise["u"], length=length, pad=pad, V=V, D=D, K=K, NB=NB,
                  embed=enc.embed_layer.weight.data, enc_weight=lin.weight.data, enc_bias=lin.bias.data,
                  category_prior=enc.category_prior, z=z, ldj=ldj, logp=logp)
    for i in range(NB):
        an, ic, cp = layers[1 + 3 * i: 4 + 3 * i]
        w, sldj = ic._get_weight("cpu")
        arrays.update({"an_bias%d" % i: an.bias.data, "an_scales%d" % i: an.scales.data,
                       "ic_w%d" % i: w, "ic_sldj%d" % i: sldj, "mask%d" % i: cp.mask,
                       "sf%d" % i: cp.scaling_factor.data, "msf%d" % i: cp.mixture_scaling_factor.data,
                       "nn_out%d" % i: captured[i],
                       "net_w0_%d" % i: cp.nn[0].weight.data, "net_b0_%d" % i: cp.nn[0].bias.data,
                       "net_w1_%d" % i: cp.nn[2].weight.data, "net_b1_%d" % i: cp.nn[2].bias.data})
    save("lm_flow_small", **arrays)


def gold_node_edge(seed):
    """GraphCNF step-2/3 block at a small Zinc-like shape: NodeEdgeFlowWrapper(ActNorm), NodeEdgeFlowWrapper(InvConv),
    NodeEdgeCoupling (experiments/molecule_generation/graph_node_edge_coupling.py) with a stand-in Edge-GNN that
    returns preset outputs; forward (training mode, regulariser on) and reverse."""
    from experiments.molecule_generation.graph_node_edge_coupling import NodeEdgeCoupling, NodeEdgeFlowWrapper
    g = torch.Generator().manual_seed(seed)
    np.random.seed(seed)
    B, N, Cn, Ce, Kn, Ke = 4, 9, 6, 2, 16, 8
    P = N * (N - 1) // 2
    length = torch.tensor([9, 5, 7, 2])
    pad = lengths_to_pad(length, N)                                        # [B,N,1]
    idx = torch.tensor([(i, j) for i in range(N) for j in range(i + 1, N)])
    mask_valid = ((idx[None, :, 0] < length[:, None]) & (idx[None, :, 1] < length[:, None])).float()   # [B,P]
    z_nodes = torch.randn(B, N, Cn, generator=g) * pad
    z_edges = torch.randn(B, P, Ce, generator=g) * mask_valid.unsqueeze(-1)
    nn_nodes = torch.randn(B, N, Cn * (2 + 3 * Kn), generator=g) * 0.6
    nn_edges = torch.randn(B, P, Ce * (2 + 3 * Ke), generator=g) * 0.6

    class _Net(nn.Module):
        def forward(self, z_nodes, z_edges, **kwargs):
            return nn_nodes, nn_edges

    cp = NodeEdgeCoupling(c_in_nodes=Cn, c_in_edges=Ce, mask_nodes=CouplingLayer.create_channel_mask(Cn),
                          mask_edges=CouplingLayer.create_channel_mask(Ce), num_mixtures_nodes=Kn, num_mixtures_edges=Ke,
                          model_func=lambda c_out_nodes, c_out_edges: _Net(), regularizer_max=3.5, regularizer_factor=2)
    cp.scaling_factor_nodes.data = torch.randn(Cn, generator=g) * 0.2
    cp.scaling_factor_edges.data = torch.randn(Ce, generator=g) * 0.2
    cp.mixture_scaling_factor_nodes.data = torch.randn(Cn, Kn, generator=g) * 0.2
    cp.mixture_scaling_factor_edges.data = torch.randn(Ce, Ke, generator=g) * 0.2
    an = NodeEdgeFlowWrapper(node_flow=ActNormFlow(Cn), edge_flow=ActNormFlow(Ce))
    ic = NodeEdgeFlowWrapper(node_flow=InvertibleConv(Cn), edge_flow=InvertibleConv(Ce))
    for f, c in ((an.node_flow, Cn), (an.edge_flow, Ce)):
        f.bias.data = torch.randn(1, 1, c, generator=g) * 0.2
        f.scales.data = torch.randn(1, 1, c, generator=g) * 0.2
    ldj0 = torch.randn(B, generator=g)
    kw = dict(length=length, channel_padding_mask=pad, mask_valid=mask_valid)
    out = {}
    with torch.no_grad():
        cp.train()
        zn, ze, ldj, detail = cp(z_nodes, z_edges, ldj=ldj0.clone(), reverse=False, **kw)
        out.update(cp_zn=zn, cp_ze=ze, cp_ldj=ldj, cp_reg_nodes=detail["regularizer_nodes_ldj"],
                   cp_reg_edges=detail["regularizer_edges_ldj"])
        cp.eval()
        zn_e, ze_e, ldj_e, _ = cp(z_nodes, z_edges, ldj=ldj0.clone(), reverse=False, **kw)
        zn_r, ze_r, ldj_r, _ = cp(zn_e, ze_e, ldj=ldj0.clone(), reverse=True, **kw)
        out.update(cp_zn_eval=zn_e, cp_ze_eval=ze_e, cp_ldj_eval=ldj_e, cp_zn_rev=zn_r, cp_ze_rev=ze_r, cp_ldj_rev=ldj_r)
        zn_a, ze_a, ldj_a = an(z_nodes, z_edges, ldj=ldj0.clone(), reverse=False, **kw)
        zn_i, ze_i, ldj_i = ic(zn_a, ze_a, ldj=ldj_a.clone(), reverse=False, **kw)
        zn_ir, ze_ir, ldj_ir = ic(zn_i, ze_i, ldj=ldj0.clone(), reverse=True, **kw)
        out.update(an_zn=zn_a, an_ze=ze_a, an_ldj=ldj_a, ic_zn=zn_i, ic_ze=ze_i, ic_ldj=ldj_i,
                   ic_zn_rev=zn_ir, ic_ze_rev=ze_ir, ic_ldj_rev=ldj_ir)
    sd = {"ic_nodes_" + k: v for k, v in ic.node_flow.state_dict().items()}
    sd.update({"ic_edges_" + k: v for k, v in ic.edge_flow.state_dict().items()})
    save("node_edge_coupling", B=B, N=N, Cn=Cn, Ce=Ce, Kn=Kn, Ke=Ke, length=length, pad=pad, mask_valid=mask_valid,
         z_nodes=z_nodes, z_edges=z_edges, nn_nodes=nn_nodes, nn_edges=nn_edges, ldj_in=ldj0,
         mask_nodes=cp.mask_nodes, mask_edges=cp.mask_edges,
         sf_nodes=cp.scaling_factor_nodes.data, sf_edges=cp.scaling_factor_edges.data,
         msf_nodes=cp.mixture_scaling_factor_nodes.data, msf_edges=cp.mixture_scaling_factor_edges.data,
         an_bias_nodes=an.node_flow.bias.data, an_scales_nodes=an.node_flow.scales.data,
         an_bias_edges=an.edge_flow.bias.data, an_scales_edges=an.edge_flow.scales.data, **sd, **out)


def _rand_graphs(g, B, N, num_edge_types, p_edge=0.2, min_len=None):
    """Random symmetric integer adjacency [B,N,N] (0 = no edge, 1..E edge type), lengths, padding zeroed."""
    length = torch.randint(min_len or max(2, N // 2), N + 1, (B,), generator=g)
    upper = (torch.rand(B, N, N, generator=g) < p_edge).long() * torch.randint(1, num_edge_types + 1, (B, N, N), generator=g)
    upper = torch.triu(upper, diagonal=1)
    adj = upper + upper.transpose(1, 2)
    valid = (torch.arange(N)[None, :] < length[:, None])
    adj = adj * (valid[:, :, None] & valid[:, None, :]).long()
    return adj, length


def _randomise(module, g, std=0.3):
    """Give every parameter a seeded non-trivial value (fresh modules have zero biases / unit LayerNorm gains)."""
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if prm.dim() >= 2:
                prm.copy_(torch.randn(prm.shape, generator=g) * (std / max(1.0, prm.shape[-1] ** 0.5) * 3.0))
            elif "norm" in name and name.endswith("weight"):
                prm.copy_(1.0 + torch.randn(prm.shape, generator=g) * 0.1)
            else:
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.1)


def gold_rgcn(name, seed, attention, num_edges, B=3, N=9, c_in=4, c_out=10, hidden=32, layers=2, skip_config=2, max_neighbours=4):
    """RGCNNet (layers/networks/graph_layers.py:157-235) with RelationGraphAttention or RelationGraphConv layers."""
    from layers.networks.graph_layers import RGCNNet, RelationGraphAttention, RelationGraphConv
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    net = RGCNNet(c_in=c_in, c_out=c_out, num_edges=num_edges, num_layers=layers, hidden_size=hidden, skip_config=skip_config,
                  max_neighbours=max_neighbours, rgc_layer_fun=RelationGraphAttention if attention else RelationGraphConv)
    _randomise(net, g)
    net.eval()
    adj, length = _rand_graphs(g, B, N, num_edges, p_edge=0.3)
    x = torch.randn(B, N, c_in, generator=g)
    pad = lengths_to_pad(length, N)
    with torch.no_grad():
        out = net(x, adjacency=adj)
        out_pad = net(x, adjacency=adj, channel_padding_mask=pad)
    sd = {"sd__" + k: v for k, v in net.state_dict().items()}
    save(name, x=x, adjacency=adj, length=length, pad=pad, out=out, out_pad=out_pad, attention=int(attention),
         num_edges=num_edges, c_in=c_in, c_out=c_out, hidden=hidden, layers=layers, skip_config=skip_config,
         max_neighbours=max_neighbours, **sd)


def gold_graph_node_flow(seed):
    """BASELINE config 3 in small: the reference's GraphNodeFlow (experiments/graph_coloring/graph_node_flow.py) -
    encoding (3 colours, d=2) + 2 x [ActNorm, InvConv, MixtureCDFCoupling(RGCNNet attention)] + ActNorm - forward
    (log-likelihood direction) in eval mode and the reverse pass of the flow layers on the resulting latents."""
    from experiments.graph_coloring.graph_node_flow import GraphNodeFlow
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)

    class _Dataset:
        @staticmethod
        def num_node_types():
            return 3

    params = {"categ_encoding": {"use_dequantization": False, "use_variational": False, "use_decoder": False,
                                 "num_dimensions": 2, "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128},
                                 "decoder_config": {"num_layers": 1, "hidden_size": 64}},
              "coupling_num_flows": 2, "coupling_hidden_size": 32, "coupling_hidden_layers": 2, "coupling_num_mixtures": 8,
              "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0}
    model = _quiet(GraphNodeFlow, params, _Dataset)
    _randomise(model, g)
    model.eval()
    B, N = 4, 8
    adj, length = _rand_graphs(g, B, N, 1, p_edge=0.35, min_len=4)
    x = torch.randint(0, 3, (B, N), generator=g)
    noise = {}

    def rec_sample(sample_shape=torch.Size()):
        noise["u"] = torch.rand(sample_shape, generator=g)
        return noise["u"]

    model.node_embed_flow.prior_distribution.distribution.sample = rec_sample
    with torch.no_grad():
        z, ldj = model(x, adjacency=adj, length=length)
        # reverse pass of the continuous layers only (decoding = argmax, tested separately)
        kw = dict(adjacency=adj, length=length, channel_padding_mask=lengths_to_pad(length, N))
        z_rev, ldj_rev = z, torch.zeros(B)
        for layer in reversed(list(model.flow_layers)[1:]):
            res = layer(z_rev, reverse=True, **kw)
            z_rev, ldj_rev = res[0], ldj_rev + res[1]      # FlowModel.forward: no ldj passed, layer ldj added (:30-44)
        x_dec = model.node_embed_flow(z_rev, reverse=True, **kw)[0]
    sd = {"sd__" + k: v for k, v in model.state_dict().items()}
    save("graph_node_flow", x=x, adjacency=adj, length=length, u=noise["u"], z=z, ldj=ldj, z_rev=z_rev, ldj_rev=ldj_rev,
         x_dec=x_dec, **sd)


def _graph_layers_with_floor_division():
    """The reference's sparse Edge-GNN path computes pair indices with `/ 2` on LongTensors (graph_layers.py:527,668), which
    torch >= 1.6 turns into a float tensor that index_select rejects.  Load the module from its source with that one
    operator restored to the integer division it was written for; nothing else is touched."""
    import importlib.util
    path = os.path.join(REF, "layers", "networks", "graph_layers.py")
    src = open(path).read()
    assert src.count("* edge_indices[...,0]) / 2 +") == 2
    src = src.replace("* edge_indices[...,0]) / 2 +", "* edge_indices[...,0]) // 2 +")
    mod = types.ModuleType("graph_layers_floor_div")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def _pairs(N, length):
    i1 = torch.tensor([i for i in range(N) for j in range(i + 1, N)])
    i2 = torch.tensor([j for i in range(N) for j in range(i + 1, N)])
    mask_valid = ((i1[None, :] < length[:, None]) & (i2[None, :] < length[:, None])).float()
    return (i1, i2), mask_valid


def gold_edge_gnn(name, seed, qkv, sparse, B=3, N=7, c_in_nodes=6, c_in_edges=2, hn=32, he=16, layers=2, max_neighbours=4):
    """EdgeGNN (graph_layers.py:737-820) with Edge2NodeAttnLayer (GraphCNF step 2) or Edge2NodeQKVAttnLayer (step 3) layers.
    ``sparse``: valid pairs = bonds only and ``binary_adjacency`` given (sparse forward + neighbour-count embedding, as in
    step 2); else valid pairs = all pairs of real nodes, dense forward (step 3)."""
    GL = _graph_layers_with_floor_division()
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    e2n = (lambda: GL.Edge2NodeQKVAttnLayer(hidden_size_nodes=hn, hidden_size_edges=he, skip_config=2)) if qkv else \
        (lambda: GL.Edge2NodeAttnLayer(hidden_size_nodes=hn, hidden_size_edges=he, skip_config=2))
    n2e = lambda: GL.Node2EdgePlainLayer(hidden_size_nodes=hn, hidden_size_edges=he, skip_config=2)
    c_out_nodes, c_out_edges = c_in_nodes * 5, c_in_edges * 8
    net = GL.EdgeGNN(c_in_nodes=c_in_nodes, c_in_edges=c_in_edges, c_out_nodes=c_out_nodes, c_out_edges=c_out_edges,
                     edge_gnn_layer_func=lambda: GL.EdgeGNNLayer(edge2node_layer_func=e2n, node2edge_layer_func=n2e),
                     max_neighbours=max_neighbours, num_layers=layers)
    _randomise(net, g)
    net.eval()
    adj, length = _rand_graphs(g, B, N, 3, p_edge=0.4, min_len=3)
    x_indices, mask_all = _pairs(N, length)
    edge_types = adj.view(B, N * N).index_select(1, x_indices[0] + x_indices[1] * N)
    mask_valid = mask_all * (edge_types != 0).float() if sparse else mask_all
    pad = lengths_to_pad(length, N)
    z_nodes = torch.randn(B, N, c_in_nodes, generator=g) * pad
    z_edges = torch.randn(B, mask_valid.shape[1], c_in_edges, generator=g) * mask_valid.unsqueeze(-1)
    binary = (adj > 0).long() if sparse else None
    with torch.no_grad():
        nodes_out, edges_out = net(z_nodes, z_edges, length=length, x_indices=x_indices, mask_valid=mask_valid,
                                   channel_padding_mask=pad, binary_adjacency=binary)
    sd = {"sd__" + k: v for k, v in net.state_dict().items()}
    save(name, z_nodes=z_nodes, z_edges=z_edges, adjacency=adj, length=length, pad=pad, mask_valid=mask_valid, x_indices1=x_indices[0],
         x_indices2=x_indices[1], nodes_out=nodes_out, edges_out=edges_out, qkv=int(qkv), sparse=int(sparse), hn=hn, he=he,
         layers=layers, max_neighbours=max_neighbours, c_in_nodes=c_in_nodes, c_in_edges=c_in_edges, c_out_nodes=c_out_nodes,
         c_out_edges=c_out_edges, **sd)


def gold_graphcnf(seed):
    """BASELINE configs 4 / 5 in small: the reference's GraphCNF (experiments/molecule_generation/graphCNF.py), forward
    (log-likelihood) and reverse (sampling) in eval mode.  Two compatibility shims so that the 2020 code runs on this torch
    (SURVEY App. B #8, #9), neither changes what is computed: integer division restored in the sparse Edge-GNN path, and the
    float64 class-prior bias of the virtual-edge decoder cast to float32."""
    sys.modules["layers.networks.graph_layers"] = _graph_layers_with_floor_division()
    from experiments.molecule_generation.graphCNF import GraphCNF
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    N, V_NODES, V_EDGES = 8, 5, 3

    class _Dataset:
        max_num_nodes = staticmethod(lambda: N)
        num_node_types = staticmethod(lambda: V_NODES)
        num_edge_types = staticmethod(lambda: V_EDGES)
        num_max_neighbours = staticmethod(lambda: 4)
        get_node_prior = staticmethod(lambda data_root="data/": np.log(np.array([0.5, 0.2, 0.15, 0.1, 0.05], dtype=np.float32)))
        get_edge_prior = staticmethod(lambda data_root="data/": np.log(np.array([0.7, 0.2, 0.1], dtype=np.float32)))

    enc = lambda d: {"use_dequantization": False, "use_variational": False, "use_decoder": False, "num_dimensions": d,
                     "flow_config": {"num_flows": 0, "hidden_layers": 2, "hidden_size": 128}, "decoder_config": {"num_layers": 1, "hidden_size": 64}}
    params = {"categ_encoding_nodes": enc(6), "categ_encoding_edges": enc(2), "coupling_hidden_size_nodes": 32, "coupling_hidden_size_edges": 16,
              "coupling_num_flows": "1,2,2", "coupling_hidden_layers": 2, "coupling_num_mixtures_nodes": 8, "coupling_num_mixtures_edges": 4,
              "coupling_mask_ratio": 0.5, "coupling_dropout": 0.0, "encoding_virtual_num_flows": 0}
    model = _quiet(GraphCNF, params, _Dataset)
    _randomise(model, g, std=0.2)
    bias = model.edge_virtual_decoder.layers.main_net[-1].bias
    bias.data = bias.data.float()
    model.eval()
    B = 4
    adj, length = _rand_graphs(g, B, N, V_EDGES, p_edge=0.3, min_len=4)
    x = torch.randint(0, V_NODES, (B, N), generator=g) * (torch.arange(N)[None, :] < length[:, None]).long()
    noise = []

    def recorder(sample_shape=torch.Size()):
        noise.append(torch.rand(sample_shape, generator=g))
        return noise[-1]

    for e in (model.node_encoding, model.edge_attr_encoding, model.edge_virtual_encoding):
        e.prior_distribution.distribution.sample = recorder
    with torch.no_grad():
        z, ldj = model(x, adjacency=adj, length=length)
    u_nodes, u_edges, u_virtual = noise
    # reverse: sample with recorded edge latents
    z_edges_init = torch.randn(B, N * (N - 1) // 2, 2, generator=g)
    z_nodes_init = torch.randn(B, N, 6, generator=g) * lengths_to_pad(length, N)
    model.prior_distribution.sample = lambda shape=None, temp=1.0, **kw: z_edges_init
    with torch.no_grad():
        (x_smp, adj_smp), ldj_smp = model(z_nodes_init, reverse=True, length=length)
    sd = {"sd__" + k: v for k, v in model.state_dict().items()}
    save("graphcnf_small", x=x, adjacency=adj, length=length, u_nodes=u_nodes, u_edges=u_edges, u_virtual=u_virtual, z=z, ldj=ldj,
         z_edges_init=z_edges_init, z_nodes_init=z_nodes_init, x_smp=x_smp, adj_smp=adj_smp, ldj_smp=ldj_smp, N=N, **sd)
    # training step of the reference (general/train.py:148-152 calls loss.backward() on the negative log-likelihood): the same
    # forward on the recorded noise with autograd enabled, gradients of every parameter -> graphcnf_small_grads.npz
    replay = iter([u_nodes, u_edges, u_virtual])
    for e in (model.node_encoding, model.edge_attr_encoding, model.edge_virtual_encoding):
        e.prior_distribution.distribution.sample = lambda sample_shape=torch.Size(): next(replay)
    model.zero_grad()
    model.train()       # training mode: 1x1 convolutions rebuild their weight inside the graph, CDF regulariser active (:108)
    z2, ldj2 = model(x, adjacency=adj, length=length)
    wz = torch.randn(z2.shape, generator=torch.Generator().manual_seed(seed + 100))
    (-(ldj2.sum()) + (z2 * wz).sum()).backward()
    grads = {"grad__" + k: p_.grad for k, p_ in model.named_parameters() if p_.grad is not None}
    save("graphcnf_small_grads", wz=wz, z_train=z2, ldj_train=ldj2, **grads)


def gold_encoding_variants(seed):
    """Encodings beyond the plain mixture model: (1) linear flows - BASELINE config 1's "4 affine couplings":
    LinearCategoricalEncoding(num_flows=4) = 4 x [ExtActNorm, InvertibleConv, affine CouplingLayer(LinearNet)]
    (linear_encoding.py:224-256), forward in eval mode and reverse decode; (2) DecoderLinear (decoder.py:35-63), the posterior
    network of the decoder-based / variational encodings.  The encodings that USE the decoder cannot run upstream:
    _decoder_forward gathers a [B*S,1,V] tensor with a [B*S,1] index (linear_encoding.py:149, variational_encoding.py:127), so
    there is nothing to record for them beyond the decoder itself.  State dicts under the reference's parameter names."""
    from layers.categorical_encoding.decoder import DecoderLinear
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    B, S, V, D = 6, 5, 4, 4
    x = torch.randint(0, V, (B, S), generator=g)
    length = torch.tensor([5, 3, 5, 1, 4, 2])
    pad = lengths_to_pad(length, S)
    out = dict(x=x, pad=pad, V=V, D=D)
    enc = _quiet(LinearCategoricalEncoding, num_dimensions=D, flow_config={"num_flows": 4, "hidden_layers": 2, "hidden_size": 32},
                 vocab_size=V)
    _randomise(enc, g, std=0.3)
    enc.eval()
    noise = {}

    def rec(sample_shape=torch.Size()):
        noise["u"] = torch.rand(sample_shape, generator=g)
        return noise["u"]

    enc.prior_distribution.distribution.sample = rec
    with torch.no_grad():
        z, ldj, _ = enc(x, reverse=False, channel_padding_mask=pad, beta=0.8)
        x_dec = enc(z, reverse=True, channel_padding_mask=pad)[0]
    out.update(flows_u=noise["u"], flows_z=z, flows_ldj=ldj, flows_x_dec=x_dec)
    out.update({"sd_flows__" + k: v for k, v in enc.state_dict().items()})
    dec = DecoderLinear(num_categories=V, embed_dim=D, hidden_size=24, num_layers=2, class_prior_log=np.log(np.array([0.4, 0.3, 0.2, 0.1], dtype=np.float32)))
    _randomise(dec, g, std=0.3)
    z_dec = torch.randn(B * S, 1, D, generator=g)
    with torch.no_grad():
        out.update(dec_z=z_dec, dec_log_probs=dec(z_dec))
    out.update({"sd_dec__" + k: v for k, v in dec.state_dict().items()})
    save("encoding_variants", **out)


def gold_dequantization(seed):
    """SigmoidFlow (sigmoid_layer.py) in both directions, with and without ldj summation, incl. inputs at
    the ends of [0,1] and far tails; VariationalDequantization (variational_dequantization.py) with 4 flows and the network of
    the reference's own usage example (:118-131), forward on recorded noise and reverse."""
    from layers.flows.sigmoid_layer import SigmoidFlow
    from layers.categorical_encoding.variational_dequantization import VariationalDequantization
    g = torch.Generator().manual_seed(seed)
    out = {}
    zs = torch.randn(4, 7, 3, generator=g) * 4.0
    zs[0, 0, 0], zs[0, 0, 1], zs[1, 2, 0] = 30.0, -30.0, 0.0
    zu = torch.rand(4, 7, 3, generator=g)
    zu[0, 0, 0], zu[0, 0, 1], zu[1, 1, 1], zu[2, 0, 0] = 0.0, 1.0, 0.5, 1.0 - 2.0 ** -20
    ldj0 = torch.randn(4, generator=g)
    with torch.no_grad():
        s_z, s_ldj = SigmoidFlow()(zs, ldj=ldj0.clone())
        s_z2, s_elem = SigmoidFlow(reverse=True)(zs, reverse=True, sum_ldj=False)
        l_z, l_ldj = SigmoidFlow()(zu, ldj=ldj0.clone(), reverse=True)
        l_z2, l_elem = SigmoidFlow(reverse=True)(zu, sum_ldj=False)
    assert torch.equal(s_z, s_z2) and torch.equal(l_z, l_z2)
    out.update(sig_in=zs, logit_in=zu, ldj0=ldj0, sig_z=s_z, sig_ldj=s_ldj, sig_elem=s_elem, logit_z=l_z, logit_ldj=l_ldj,
               logit_elem=l_elem)

    B, S, V, E, H, NF = 5, 7, 6, 12, 20, 4

    class ExampleNetwork(nn.Module):        # as in the reference's example (variational_dequantization.py:118-131)
        def __init__(self, c_out):
            super().__init__()
            self.inp_layer = nn.Linear(1, H)
            self.main_net = nn.Sequential(nn.Linear(H + E, H), nn.ReLU(), nn.Linear(H, c_out))

        def forward(self, x, ext_input, **kwargs):
            return self.main_net(torch.cat([self.inp_layer(x), ext_input], dim=-1))

    torch.manual_seed(seed)
    deq = VariationalDequantization(flow_config={"num_flows": NF, "model_func": lambda c_out: ExampleNetwork(c_out),
                                                 "block_type": "Linear"}, vocab_size=V, default_embed_layer_dims=E)
    _randomise(deq, g, std=0.3)
    deq.eval()
    x = torch.randint(0, V, (B, S), generator=g)
    torch.manual_seed(seed + 1)
    u = torch.rand_like(x, dtype=torch.float32)          # the draw forward() makes first (:39)
    torch.manual_seed(seed + 1)
    with torch.no_grad():
        z_cont, ldj = deq(x, reverse=False)
        x_rec, _ = deq(z_cont, reverse=True)
    assert torch.equal(x_rec, x)
    out.update(x=x, u=u, z_cont=z_cont, ldj=ldj, x_rec=x_rec, V=V, num_flows=NF)
    out.update({"sd__" + k: v for k, v in deq.state_dict().items()})
    save("dequantization", **out)


if __name__ == "__main__":
    torch.set_num_threads(4)
    if ONLY:
        for _n in ONLY:
            {"node_edge": lambda: gold_node_edge(seed=22),
             "rgcn": lambda: (gold_rgcn("rgcn_attention", 23, True, 1), gold_rgcn("rgcn_attention_e3", 24, True, 3, skip_config=1),
                              gold_rgcn("rgcn_conv", 25, False, 3, N=12), gold_rgcn("rgcn_conv_skip0", 26, False, 1, skip_config=0,
                                                                                   max_neighbours=0)),
             "graph_flow": lambda: gold_graph_node_flow(seed=27),
             "graphcnf": lambda: gold_graphcnf(seed=32),
             "encoding_variants": lambda: gold_encoding_variants(seed=33),
             "dequantization": lambda: gold_dequantization(seed=34),
             "edge_gnn": lambda: (gold_edge_gnn("edge_gnn_attn_sparse", 28, False, True), gold_edge_gnn("edge_gnn_attn_dense", 29, False, False),
                                  gold_edge_gnn("edge_gnn_qkv_dense", 30, True, False, N=9), gold_edge_gnn("edge_gnn_qkv_sparse", 31, True, True))}[_n]()
        sys.exit(0)
    gold_mixcdf_selftest()
    gold_mixcdf("mixcdf_lm_small", 3, 32, 16, 8, seed=1)
    gold_mixcdf("mixcdf_lm_padded_sf", 4, 40, 16, 8, seed=2, padded=True, sf_std=0.3)
    gold_mixcdf("mixcdf_stress", 2, 16, 16, 8, seed=3, nn_std=2.0, z_std=2.0)
    gold_mixcdf("mixcdf_mol_nodes", 5, 38, 6, 16, seed=4, padded=True, sf_std=0.2, reg_max=3.5, reg_factor=2.0)
    gold_mixcdf("mixcdf_mol_edges", 3, 71, 2, 8, seed=5, padded=True, reg_max=3.5, training=False)
    gold_mixcdf("mixcdf_chess", 4, 9, 1, 4, seed=6, chess=True, padded=True)
    gold_mixcdf("mixcdf_chess_flip", 4, 10, 1, 4, seed=7, chess=True, flip=True)
    gold_mixcdf("mixcdf_flip_k10", 3, 6, 4, 10, seed=8, flip=True, sf_std=0.4)
    gold_mixcdf("mixcdf_ratio_k3", 2, 5, 5, 3, seed=9, ratio=0.3, z_std=3.0)
    gold_mixcdf_tails("mixcdf_tails", seed=10)
    gold_mixcdf_tails("mixcdf_right_tail", seed=10, right_tail=True)
    gold_mixcdf_inv_underflow()
    gold_autoregressive(seed=11)
    gold_affine(seed=12)
    gold_actnorm(seed=13)
    gold_ext_actnorm(seed=14)
    gold_invconv(seed=15)
    gold_logistic(seed=16)
    gold_encoding("encode_lm", 3, 20, 51, 16, seed=17)
    gold_encoding("encode_mol_nodes", 4, 38, 9, 6, seed=18, padded=True, beta=0.7, prior_std=1.0)
    gold_encoding("encode_mol_edges", 3, 50, 3, 2, seed=19, padded=True, prior_std=0.5, training=True)
    gold_encoding("encode_virtual", 2, 12, 1, 2, seed=20)
    gold_lm_flow(seed=21)
    gold_node_edge(seed=22)
    gold_rgcn("rgcn_attention", 23, True, 1)
    gold_rgcn("rgcn_attention_e3", 24, True, 3, skip_config=1)
    gold_rgcn("rgcn_conv", 25, False, 3, N=12)
    gold_rgcn("rgcn_conv_skip0", 26, False, 1, skip_config=0, max_neighbours=0)
    gold_graph_node_flow(seed=27)
    gold_edge_gnn("edge_gnn_attn_sparse", 28, False, True)
    gold_edge_gnn("edge_gnn_attn_dense", 29, False, False)
    gold_edge_gnn("edge_gnn_qkv_dense", 30, True, False, N=9)
    gold_edge_gnn("edge_gnn_qkv_sparse", 31, True, True)
    gold_graphcnf(seed=32)
    gold_encoding_variants(seed=33)
    gold_dequantization(seed=34)
