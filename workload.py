"""Synthetic workloads of BASELINE.json's configs, shared by bench.py, __graft_entry__.smoke() and
tests/ (not part of the product package).

LM workload (configs[1], SURVEY.md section 8d "Config 2"): tokens [B, S=256] over V=51 classes,
mixture-of-logistics encoding into d=16 latents, then 8 flow blocks
[ActNorm, InvertibleConv 1x1, MixtureCDFCoupling K=8, channel mask ratio 0.5], logistic prior.

The coupling network is a black box to the hot path (coupling_layer.py:28-39).  Two forms:
  * "params given": the per-coupling network outputs nn_out [B,S,C*(2+3K)] are inputs resident in
    HBM (kernel-level view of section 8d);
  * "stand-in net": one per-position Linear(C -> C*(2+3K)) on the masked latents, so that the
    module-level API (FlowModel of drop-in layers) can be driven end to end from tokens.
All parameters are generated on the CPU from one seed so the GPU path, the CPU oracle and the
world_size>1 shards see identical values.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

LM = dict(B=4096, S=256, D=16, K=8, V=51, blocks=8, embed=64)


@dataclass
class LMParams:
    S: int
    D: int
    K: int
    V: int
    embed_w: torch.Tensor      # [V, E]
    pred_w: torch.Tensor       # [2D, E]
    pred_b: torch.Tensor       # [2D]
    prior: torch.Tensor        # [V] log-softmaxed
    blocks: list               # dicts: bias, scales [C]; p, l, u [C,C]; log_s, sign_s [C]; net_w [C*(2+3K), C]; net_b; sf; msf

    def table(self):
        return torch.nn.functional.linear(self.embed_w, self.pred_w, self.pred_b)


def _lu_factors(rng, C):
    """Random rotation -> (P, L, U, log|s|, sign s) like InvertibleConv.__init__ (permutation_layers.py:33-54)."""
    import scipy.linalg
    q = np.linalg.qr(rng.standard_normal((C, C)))[0].astype(np.float32)
    p, l, u = scipy.linalg.lu(q)
    s = np.diag(u)
    f = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return f(p), f(l), f(np.triu(u, k=1)), f(np.log(np.abs(s))), f(np.sign(s))


def lm_params(seed=0, S=LM["S"], D=LM["D"], K=LM["K"], V=LM["V"], blocks=LM["blocks"], embed=LM["embed"]) -> LMParams:
    g = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    rn = lambda *shape, std=1.0: torch.randn(*shape, generator=g) * std
    pn = 2 + 3 * K
    embed_w = rn(V, embed)
    pred_w = rn(2 * D, embed, std=1.0 / math.sqrt(embed))
    pred_w[D:] *= 0.3                      # raw log-scales: modest
    pred_b = rn(2 * D, std=0.1)
    blks = []
    for _ in range(blocks):
        p, l, u, log_s, sign_s = _lu_factors(rng, D)
        blks.append(dict(
            bias=rn(D, std=0.1), scales=rn(D, std=0.1), p=p, l=l, u=u, log_s=log_s + rn(D, std=0.05), sign_s=sign_s,
            # stand-in coupling net: nn_out ~ N(0, 0.5^2) for unit-scale latents (section 8d)
            net_w=rn(D * pn, D, std=0.5 / math.sqrt(D / 2)), net_b=rn(D * pn, std=0.1),
            sf=rn(D, std=0.3), msf=rn(D, K, std=0.3),
            mask=torch.cat([torch.ones(D // 2), torch.zeros(D - D // 2)]).view(1, D)))
    return LMParams(S=S, D=D, K=K, V=V, embed_w=embed_w, pred_w=pred_w, pred_b=pred_b,
                    prior=torch.log_softmax(rn(V, std=0.5), dim=-1), blocks=blks)


INIT_B, INIT_S = 64, 32   # batch used for the data-dependent ActNorm initialisation


def data_init_oracle(prm, seed=0):
    """CPU arm: set every block's ActNorm (bias, scales) by the reference's data-dependent
    initialisation (flow_model.py:95-131, activation_normalization.py:55-67) on a small seeded
    batch, so latents stay unit-scale through the stack like in an initialised reference model.
    The GPU arm does the same with its own kernels (LMDevicePath.data_init)."""
    from oracle import cnf_oracle as O
    Bc, Sc = INIT_B, min(prm.S, INIT_S)
    tokens, u = lm_tokens(Bc, Sc, prm.V, seed=seed + 7), lm_noise(Bc, Sc, prm.D, seed=seed + 7)
    z, _, _ = O.categ_encode(tokens, u, prm.table(), prm.prior)
    for b in prm.blocks:
        bias, scales = O.actnorm_data_init(z)
        b["bias"], b["scales"] = bias.flatten().clone(), scales.flatten().clone()
        z, _ = O.actnorm(z, bias, scales)
        w, sldj = O.invconv_weight(b["p"], b["l"], b["log_s"], b["u"], b["sign_s"])
        z, _ = O.invconv(z, w, sldj)
        m = O.expand_mask(b["mask"], z)
        nn_out = torch.nn.functional.linear(z * m, b["net_w"], b["net_b"])
        z, _, _ = O.mixcdf_coupling(z, nn_out, m, prm.K, b["sf"], b["msf"], training=False)
    return prm


def lm_tokens(B, S, V, seed=0):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randint(0, V, (B, S), generator=g, dtype=torch.int64)


def lm_noise(B, S, D, seed=0):
    g = torch.Generator().manual_seed(2000 + seed)
    return torch.rand(B * S, 1, D, generator=g)


# --------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle's restatement of the reference composition (oracle/cnf_oracle.py)
# --------------------------------------------------------------------------------------------------
def lm_oracle_forward(prm: LMParams, tokens, u_noise, nn_outs=None):
    """(z, ldj [B], log_prior [B]) on the CPU through oracle.lm_flow_forward.  ``nn_outs`` (list of
    [B,S,C*(2+3K)]) selects the params-given form, otherwise the stand-in net is evaluated."""
    from oracle import cnf_oracle as O
    blocks = []
    for i, b in enumerate(prm.blocks):
        w, sldj = O.invconv_weight(b["p"], b["l"], b["log_s"], b["u"], b["sign_s"])
        blk = dict(bias=b["bias"].view(1, 1, -1), scales=b["scales"].view(1, 1, -1), weight=w, sldj=sldj,
                   mask=b["mask"], K=prm.K, sf=b["sf"], msf=b["msf"])
        if nn_outs is not None:
            blk["nn_out"] = nn_outs[i]
        else:
            blk["nn_fn"] = (lambda zin, b=b: torch.nn.functional.linear(zin, b["net_w"], b["net_b"]))
        blocks.append(blk)
    return O.lm_flow_forward(tokens, u_noise, dict(table=prm.table(), prior=prm.prior), blocks)


# --------------------------------------------------------------------------------------------------
# GPU arm, kernel-level: C-ABI ops with the coupling parameters given
# --------------------------------------------------------------------------------------------------
class LMDevicePath:
    """encode -> blocks x [actnorm, 1x1 conv, mixture coupling] -> prior log-prob, all through
    categoricalnf_b200.ops (one C-ABI launch each)."""

    def __init__(self, prm: LMParams, device):
        from categoricalnf_b200 import ops
        self.ops, self.prm, self.dev = ops, prm, device
        d = lambda t: t.to(device).contiguous()
        self.table, self.prior = d(prm.table()), d(prm.prior)
        self.blocks = []
        for b in prm.blocks:
            w, _, sldj = ops.invconv_build(p=d(b["p"]), l=d(b["l"]), u=d(b["u"]), log_s=d(b["log_s"]), sign_s=d(b["sign_s"]),
                                           want_inverse=False)
            self.blocks.append(dict(bias=d(b["bias"]), scales=d(b["scales"]), w=w, sldj=sldj, sf=d(b["sf"]), msf=d(b["msf"]),
                                    mask_c=b["mask"].flatten().tolist(), net_w=d(b["net_w"]), net_b=d(b["net_b"])))
        self.mix_events = []   # (start, end) CUDA events around every mixture-coupling launch when timing
        self.ldj_const = None

    def data_init(self, seed=0):
        """Data-dependent ActNorm initialisation with the product's own kernels
        (cnf_actnorm_data_init), block by block on a small seeded batch; the resulting (bias,
        scales) are also written back into ``prm`` so that the CPU oracle arm can replay them."""
        ops, prm = self.ops, self.prm
        Bc, Sc = INIT_B, min(prm.S, INIT_S)
        tokens = lm_tokens(Bc, Sc, prm.V, seed=seed + 7).to(self.dev)
        u = lm_noise(Bc, Sc, prm.D, seed=seed + 7).to(self.dev)
        ldj = torch.zeros(Bc, dtype=torch.float32, device=self.dev)
        z, ldj, _ = ops.categ_encode(tokens, self.table, self.prior, ldj, noise=u)
        for b, pb in zip(self.blocks, prm.blocks):
            b["bias"], b["scales"] = ops.actnorm_data_init(z)
            pb["bias"], pb["scales"] = b["bias"].cpu(), b["scales"].cpu()
            self.ldj_const = None
            z, _ = ops.actnorm(z, b["bias"], b["scales"], None)
            z, _ = ops.invconv_apply(z, b["w"], b["sldj"], None)
            zin = z * torch.tensor(b["mask_c"], device=self.dev)
            nn_out = ops.linear(zin, b["net_w"], b["net_b"])      # the product's tcgen05 projection (3xTF32), not cuBLAS
            z, _, _ = ops.mixcdf(z, nn_out, prm.K, mask_c=b["mask_c"], scaling_factor=b["sf"],
                                 mixture_scaling_factor=b["msf"])
        return self

    def forward(self, tokens, nn_outs=None, u_noise=None, seed=0, offset=0, time_mix=False, fused=True, total=None):
        """-> (z, ldj [B], log_prior [B]).  ``nn_outs`` None evaluates the stand-in net with torch.
        ``total`` (float64 [2] on the device): the last kernel - the prior log-prob - also adds ldj, so the third return
        value is the per-sample LOG-LIKELIHOOD ldj + log_prior, and leaves (sum log-likelihood, B) in ``total``.
        ``fused``: ActNorm + 1x1 conv of block i+1 run in the epilogue of the kernel producing its
        input (encode for block 0, mixture coupling i otherwise); identical results, 2 launches and
        two passes over z fewer per block."""
        ops, prm = self.ops, self.prm
        B, S = tokens.shape
        nb = len(self.blocks)
        ldj = torch.zeros(B, dtype=torch.float32, device=self.dev)
        fused = fused and ops.categ_encode_fusable(B, S, prm.V, prm.D)

        def nxt(i):
            b = self.blocks[i]
            return (b["bias"], b["scales"], b["w"])

        if fused:
            if self.ldj_const is None:   # sum over blocks of (sum scales + sldj): per-sample constant x S
                self.ldj_const = torch.stack([b["scales"].sum() + b["sldj"].reshape(()) for b in self.blocks]).sum().reshape(1)
            z, ldj, _ = ops.categ_encode(tokens, self.table, self.prior, ldj, noise=u_noise, seed=seed, offset=offset,
                                         fuse_next=nxt(0))
        else:
            z, ldj, _ = ops.categ_encode(tokens, self.table, self.prior, ldj, noise=u_noise, seed=seed, offset=offset)
        for i, b in enumerate(self.blocks):
            if not fused:
                z, ldj = ops.actnorm(z, b["bias"], b["scales"], ldj)
                z, ldj = ops.invconv_apply(z, b["w"], b["sldj"], ldj)
            if nn_outs is not None:
                nn_out = nn_outs[i]
            else:
                zin = z * torch.tensor(b["mask_c"], device=self.dev)
                nn_out = ops.linear(zin, b["net_w"], b["net_b"])      # the product's tcgen05 projection (3xTF32), not cuBLAS
            if time_mix:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            fuse_next = nxt(i + 1) if (fused and i + 1 < nb) else None
            if fuse_next is not None and not ops.mixcdf_fusable(z, nn_out, prm.K, mask_c=b["mask_c"]):
                raise RuntimeError("LMDevicePath: coupling %d is not fusable although the encode was" % i)
            z, ldj, _ = ops.mixcdf(z, nn_out, prm.K, mask_c=b["mask_c"], scaling_factor=b["sf"],
                                   mixture_scaling_factor=b["msf"], ldj=ldj, fuse_next=fuse_next)
            if time_mix:
                e1.record()
                self.mix_events.append((e0, e1))
        if fused:
            ops.ldj_axpy(ldj, alpha=float(S), alpha_dev=self.ldj_const)
        if total is not None:
            ll, _ = ops.logistic_logprob(z, add=ldj, total=total)
            return z, ldj, ll
        logp, _ = ops.logistic_logprob(z)
        return z, ldj, logp


def bits_per_dim(ldj, logp, S):
    return float(((-ldj.double() - logp.double()) / S).mean() * math.log2(math.e))


# --------------------------------------------------------------------------------------------------
# GPU arm, module-level: the drop-in FlowModel a user of the reference would build
# --------------------------------------------------------------------------------------------------
class StandInNet(torch.nn.Module):
    """Per-position Linear coupling network with the call signature coupling_layer.py:28-35 uses.  The
    projection is the product's tcgen05 Linear (a16); the network also exposes the final-projection
    protocol (categoricalnf_b200.layers.networks.split_final_linear) so that the coupling layer can fuse
    it with the transform at evaluation time."""

    def __init__(self, c_in, c_out):
        super().__init__()
        from categoricalnf_b200.layers.networks import TCLinear
        self.lin = TCLinear(c_in, c_out)

    cnf_features_are_input = True      # the network IS its final Linear: the coupling may fold its mask into the weight

    @property
    def cnf_final_linear(self):
        return self.lin

    def cnf_features(self, x, length=None, **kwargs):
        return x

    def forward(self, x, length=None, **kwargs):
        return self.lin(x)


def _lm_model_from(prm: LMParams, C, net_factory):
    """The LM flow built from the classes in namespace ``C`` (the drop-in modules or the reference's own - they share
    constructor signatures and parameter names) with ``prm`` loaded.  -> (FlowModel, prior)."""
    import contextlib
    import io
    D, K = prm.D, prm.K
    with contextlib.redirect_stdout(io.StringIO()):
        enc = C.LinearCategoricalEncoding(num_dimensions=D, flow_config={"num_flows": 0, "hidden_layers": 2, "hidden_size": 64},
                                          vocab_size=prm.V, default_embed_layer_dims=prm.embed_w.shape[1],
                                          category_prior=prm.prior.clone())
    layers = [enc]
    with torch.no_grad():
        enc.embed_layer.weight.copy_(prm.embed_w)
        enc.flow_layers[0].pred_net.layer.weight.copy_(prm.pred_w)
        enc.flow_layers[0].pred_net.layer.bias.copy_(prm.pred_b)
        for b in prm.blocks:
            an = C.ActNormFlow(c_in=D, data_init=False)
            an.bias.copy_(b["bias"].view(1, 1, -1))
            an.scales.copy_(b["scales"].view(1, 1, -1))
            conv = C.InvertibleConv(c_in=D)
            conv.p.copy_(b["p"]); conv.l.copy_(b["l"]); conv.u.copy_(b["u"])
            conv.log_s.copy_(b["log_s"]); conv.sign_s.copy_(b["sign_s"])
            mix = C.MixtureCDFCoupling(c_in=D, mask=b["mask"].clone(), model_func=lambda c_out: net_factory(D, c_out),
                                       num_mixtures=K)
            mix.scaling_factor.copy_(b["sf"]); mix.mixture_scaling_factor.copy_(b["msf"])
            mix.nn.lin.weight.copy_(b["net_w"]); mix.nn.lin.bias.copy_(b["net_b"])
            layers += [an, conv, mix]
    with contextlib.redirect_stdout(io.StringIO()):
        model = C.FlowModel(layers, name="LM flow (synthetic)")
    return model, C.LogisticDistribution(mu=0.0, sigma=1.0)


def build_lm_model(prm: LMParams, device):
    """(encoding + blocks as a drop-in FlowModel, prior) with ``prm`` loaded, in eval mode on ``device``."""
    from types import SimpleNamespace
    from categoricalnf_b200.layers.categorical_encoding import LinearCategoricalEncoding
    from categoricalnf_b200.layers.flows import (ActNormFlow, FlowModel, InvertibleConv, LogisticDistribution,
                                                 MixtureCDFCoupling)
    C = SimpleNamespace(LinearCategoricalEncoding=LinearCategoricalEncoding, ActNormFlow=ActNormFlow, FlowModel=FlowModel,
                        InvertibleConv=InvertibleConv, LogisticDistribution=LogisticDistribution,
                        MixtureCDFCoupling=MixtureCDFCoupling)
    model, prior = _lm_model_from(prm, C, StandInNet)
    return model.to(device).eval(), prior.to(device)


# --------------------------------------------------------------------------------------------------
# CPU reference arm proper: the UNMODIFIED reference modules from baseline/_ref (tools/vendor_reference.sh)
# --------------------------------------------------------------------------------------------------
import os as _os

REF_ROOT = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "baseline", "_ref")


def reference_available() -> bool:
    return _os.path.isdir(_os.path.join(REF_ROOT, "layers", "flows"))


def import_reference():
    """Namespace of the reference's own flow classes, imported from baseline/_ref.  Must not be mixed with
    ``categoricalnf_b200.install`` in one process (both claim the top-level ``layers`` package)."""
    import sys
    from types import SimpleNamespace
    if "categoricalnf_b200.install" in sys.modules and any(
            getattr(sys.modules.get(n), "__name__", "").startswith("categoricalnf_b200.") for n in ("layers.flows.flow_model",)):
        raise RuntimeError("import_reference: the drop-in modules are installed under the reference's names in this process")
    if not reference_available():
        raise FileNotFoundError("baseline/_ref is missing: run tools/vendor_reference.sh where /root/reference exists")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    try:
        import matplotlib  # noqa: F401
    except ImportError:      # general/task.py:8 and layers/flows/flow_model.py import it without using it here
        import types
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pyplot, mpl.colors = types.ModuleType("matplotlib.pyplot"), types.ModuleType("matplotlib.colors")
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": mpl.pyplot, "matplotlib.colors": mpl.colors})
    from layers.categorical_encoding.linear_encoding import LinearCategoricalEncoding
    from layers.flows.activation_normalization import ActNormFlow
    from layers.flows.distributions import LogisticDistribution
    from layers.flows.flow_model import FlowModel
    from layers.flows.mixture_cdf_layer import MixtureCDFCoupling
    from layers.flows.permutation_layers import InvertibleConv
    assert FlowModel.__module__ == "layers.flows.flow_model" and REF_ROOT in _os.path.abspath(sys.modules[FlowModel.__module__].__file__)
    return SimpleNamespace(LinearCategoricalEncoding=LinearCategoricalEncoding, ActNormFlow=ActNormFlow, FlowModel=FlowModel,
                           InvertibleConv=InvertibleConv, LogisticDistribution=LogisticDistribution,
                           MixtureCDFCoupling=MixtureCDFCoupling)


class _RefStandInNet(torch.nn.Module):
    """The stand-in per-position Linear coupling network for the reference arm: plain ``nn.Linear`` (the reference's
    networks are plain torch modules), same call signature (coupling_layer.py:28-35)."""

    def __init__(self, c_in, c_out):
        super().__init__()
        self.lin = torch.nn.Linear(c_in, c_out)

    def forward(self, x, length=None, **kwargs):
        return self.lin(x)


def build_lm_reference_model(prm: LMParams):
    """The same LM flow from the reference's own classes on the CPU, eval mode.  -> (FlowModel, prior)."""
    model, prior = _lm_model_from(prm, import_reference(), _RefStandInNet)
    return model.eval(), prior


def lm_reference_forward(model, prior, tokens, seed):
    """(z, ldj [B], log_prior [B], u_noise) through the unmodified reference on the CPU.  The encoding draws its noise
    from torch's global CPU generator (linear_encoding.py:78): seeding it makes the draw reproducible, and the same draw
    is returned so the oracle port / the GPU path can be run on identical noise."""
    B, S = tokens.shape
    D = model.flow_layers[0].D
    torch.manual_seed(seed)
    u = torch.rand(B * S, 1, D)                # what Uniform(0,1).sample((B*S,1,D)) will draw next
    torch.manual_seed(seed)
    with torch.no_grad():
        z, ldj = model(tokens, reverse=False)
        logp = prior.log_prob(z).sum(dim=[1, 2])
    return z, ldj, logp, u
