"""CPU oracle for the coupling-layer hot path of phlippe/CategoricalNF.

TEST INFRASTRUCTURE ONLY.  Nothing in ``categoricalnf_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the CPU arm that is being compared against - never as the product path.

What it is: a restatement, in plain eager PyTorch on the CPU, of the arithmetic
the reference performs for each function on the hot path.
Like the reference it up-casts the mixture transform to float64 and rounds the
results back to float32.  Every function cites the reference ``file:line`` it
follows (paths relative to the reference checkout).

Parity pinning: the reference ships no golden vectors, so
the oracle is pinned against outputs of the reference implementation itself,
generated in the authoring container by ``tests/golden/make_golden.py`` and
committed as ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

Shapes: ``z`` is ``[B, S, C]`` float32, ``nn_out`` is ``[B, S, C*(2+3K)]`` float32
with the per-channel record ``[t, log_s, log_pi x K, mu x K, log_scale x K]``.
``mask`` is the coupling mask (1 = conditioner input, 0 = transformed) with shape
``[1, C]`` (channel mask) or ``[S_m, 1]`` (chess mask tiled along S).
``pad`` is the ``channel_padding_mask`` ``[B, S, 1]`` (1 = real token) or None.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LOGISTIC_SIGMA = 1.0 / 1.81          # layers/flows/distributions.py:94
LOGISTIC_LOG_SIGMA = float(np.log(LOGISTIC_SIGMA))
LOGISTIC_EPS = 1e-4                  # layers/flows/distributions.py:93


# ---------------------------------------------------------------------------
# masks  (layers/flows/coupling_layer.py:67-74, 101-121)
# ---------------------------------------------------------------------------
def channel_mask(c_in: int, ratio: float = 0.5, mask_floor: bool = True) -> torch.Tensor:
    """coupling_layer.py:101-112 - first floor(c_in*ratio) channels are conditioner inputs."""
    n_cond = int(math.floor(c_in * ratio)) if mask_floor else int(math.ceil(c_in * ratio))
    m = torch.zeros(1, c_in)
    m[0, :n_cond] = 1.0
    return m


def chess_mask(seq_len: int = 2) -> torch.Tensor:
    """coupling_layer.py:115-121 - ceil(seq_len/2) ones followed by zeros, shape [seq_len, 1]."""
    assert seq_len > 1
    n_zero = seq_len // 2
    m = torch.zeros(seq_len, 1)
    m[: seq_len - n_zero, 0] = 1.0
    return m


def expand_mask(mask: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    """coupling_layer.py:67-74 - broadcast the stored mask to ``[1, S|1, C|1]`` for ``z``."""
    m = mask.unsqueeze(0) if z.dim() > mask.dim() else mask
    S = z.size(1)
    if 1 < m.size(1) < S:
        m = m.repeat(1, int(math.ceil(S / m.size(1))), 1)
    if m.size(1) > S:
        m = m[:, :S]
    return m


# ---------------------------------------------------------------------------
# a1: parameter split   (layers/flows/mixture_cdf_layer.py:145-180)
# ---------------------------------------------------------------------------
def mixt_params(nn_out, mask, num_mixtures, scaling_factor=None, mixture_scaling_factor=None):
    K = num_mixtures
    pn = 2 + 3 * K
    rec = nn_out.reshape(nn_out.shape[:-1] + (nn_out.shape[-1] // pn, pn))
    t, log_s = rec[..., 0], rec[..., 1]
    log_pi, mu, mls = rec[..., 2:2 + K], rec[..., 2 + K:2 + 2 * K], rec[..., 2 + 2 * K:2 + 3 * K]
    # tanh bounding happens in float32, before the up-cast (:157-162)
    if scaling_factor is not None:
        fac = scaling_factor.exp()
        log_s = torch.tanh(log_s / fac.clamp(min=1.0)) * fac
    if mixture_scaling_factor is not None:
        mfac = mixture_scaling_factor.exp()
        mls = torch.tanh(mls / mfac.clamp(min=1.0)) * mfac
    if mask is not None:                      # (:165-171)
        keep = 1 - mask
        t, log_s = t * keep, log_s * keep
        keep_k = keep.unsqueeze(-1)
        log_pi, mu, mls = log_pi * keep_k, mu * keep_k, mls * keep_k
    return t.double(), log_s.double(), log_pi.double(), mu.double(), mls.double()


def _safe_log(x):
    """mixture_cdf_layer.py:197-198"""
    return torch.log(x.clamp(min=1e-22))


def _mix_log_cdf(x, log_pi, mu, mls):
    """mixture_cdf_layer.py:209-214, 226-232"""
    u = (x.unsqueeze(-1) - mu) * torch.exp(-mls)
    return torch.logsumexp(F.log_softmax(log_pi, dim=-1) + F.logsigmoid(u), dim=-1)


def _mix_log_pdf(x, log_pi, mu, mls):
    """mixture_cdf_layer.py:201-206, 217-223"""
    u = (x.unsqueeze(-1) - mu) * torch.exp(-mls)
    comp = u - mls - 2 * F.softplus(u)
    return torch.logsumexp(F.log_softmax(log_pi, dim=-1) + comp, dim=-1)


def _bisect_inv_cdf(y, log_pi, mu, mls, eps=1e-10, max_iters=100):
    """mixture_cdf_layer.py:235-264 - global-stop bisection started at x=0."""
    if y.min() <= 0 or y.max() >= 1:
        raise RuntimeError("Inverse logisitic CDF got y outside (0, 1)")
    x = torch.zeros_like(y)
    span = torch.exp(mls).sum(dim=-1, keepdim=True)
    lb = (mu - 20 * span).min(dim=-1)[0]
    ub = (mu + 20 * span).max(dim=-1)[0]
    it, diff = 0, float("inf")
    while diff > eps and it < max_iters:
        above = (torch.exp(_mix_log_cdf(x, log_pi, mu, mls)) > y).to(y.dtype)
        below = 1 - above
        new_x = above * (x + lb) / 2.0 + below * (x + ub) / 2.0
        lb = above * lb + below * x
        ub = above * x + below * ub
        diff = (new_x - x).abs().max()
        x = new_x
        it += 1
    return x


# ---------------------------------------------------------------------------
# a2 / a3: the transform   (layers/flows/mixture_cdf_layer.py:95-142)
# ---------------------------------------------------------------------------
def mixcdf_run(z, t, log_s, log_pi, mu, mls, *, reverse=False, mask=None, pad=None,
               reg_max=-1.0, reg_factor=1.0, training=True):
    """Returns (z_out float64, ldj float64 [B], reg float64 [B,S,C] or None)."""
    x = z.double()
    change = (1 - mask) if mask is not None else torch.ones_like(x)
    if pad is not None:
        change = change * pad
    reg = None
    if not reverse:
        cdf = _mix_log_cdf(x, log_pi, mu, mls).exp()                       # :105
        if reg_max > 0 and training:                                       # :108-112
            reg = torch.stack([_safe_log(cdf), _safe_log(1 - cdf)], dim=-1) / np.log(10)
            reg = (reg.clamp(max=-reg_max) + reg_max).sum(dim=-1) * change
        else:
            reg = torch.zeros_like(cdf)                                    # :114
        y = -_safe_log(cdf.reciprocal() - 1.0)                             # :117, :273
        mixt_ldj = -_safe_log(cdf) - _safe_log(1.0 - cdf)                  # :274
        out = (y + t) * log_s.exp()                                        # :119
        log_f = _mix_log_pdf(x, log_pi, mu, mls)                           # :121
        ldj = (change * (log_s + mixt_ldj + log_f + reg * reg_factor)).sum(dim=[1, 2])
    else:
        y = x * (-log_s).exp() - t                                         # :126
        cdf = torch.sigmoid(y)                                             # :128, :269
        mixt_ldj = F.softplus(y) + F.softplus(-y)                          # :270
        cdf = cdf.clamp(1e-5, 1.0 - 1e-5)                                  # :130
        out = _bisect_inv_cdf(cdf, log_pi, mu, mls)                        # :132
        log_f = _mix_log_pdf(out, log_pi, mu, mls)                         # :134
        ldj = -(change * (log_s + mixt_ldj + log_f)).sum(dim=[1, 2])       # :136
    if mask is not None:                                                   # :137-138
        out = out * change + x * (1 - change)
    return out, ldj, reg


def mixcdf_coupling(z, nn_out, mask, num_mixtures, scaling_factor, mixture_scaling_factor, *,
                    reverse=False, pad=None, reg_max=-1.0, reg_factor=1.0, training=True):
    """``MixtureCDFCoupling.forward`` after the network call (mixture_cdf_layer.py:58-92).

    ``mask`` must already be expanded with :func:`expand_mask`.  Returns
    ``(z_out f32, ldj f32 [B], reg_ldj f32 [B])`` - the incoming ldj is ignored
    by the reference layer (App. B #1).
    """
    p = mixt_params(nn_out, mask, num_mixtures, scaling_factor, mixture_scaling_factor)
    pad_full = pad if pad is not None else torch.ones_like(z)               # :49-50
    out, ldj, reg = mixcdf_run(z, *p, reverse=reverse, mask=mask, pad=pad_full,
                               reg_max=reg_max, reg_factor=reg_factor, training=training)
    out = out.float() * pad_full                                            # :74-76
    reg_b = reg.float().sum(dim=[1, 2]) if reg is not None else torch.zeros_like(ldj).float()
    return out, ldj.float(), reg_b


def autoregressive_mixcdf(z, nn_out, num_mixtures, scaling_factor, mixture_scaling_factor,
                          ldj=None, pad=None):
    """layers/flows/autoregressive_coupling.py:25-47 - mask=None, accumulates ldj, forward only."""
    if ldj is None:
        ldj = z.new_zeros(z.size(0))
    p = mixt_params(nn_out, None, num_mixtures, scaling_factor, mixture_scaling_factor)
    out, l, _ = mixcdf_run(z, *p, reverse=False, mask=None, pad=None)
    out = out.float()
    if pad is not None:
        out = out * pad
    return out, ldj + l.float()


def node_edge_coupling(z_nodes, z_edges, nn_nodes, nn_edges, mask_nodes, mask_edges, k_nodes, k_edges,
                       sf_nodes, sf_edges, msf_nodes, msf_edges, *, ldj=None, reverse=False, pad=None, mask_valid=None,
                       reg_max=-1.0, reg_factor=1.0, training=True):
    """a14: ``NodeEdgeCoupling.forward`` after the Edge-GNN call
    (experiments/molecule_generation/graph_node_edge_coupling.py:63-110): network outputs and transformed latents
    are zeroed at padded nodes / invalid pairs, both ldj are ADDED to the incoming one (:99).
    ``pad`` is ``[B,N,1]``, ``mask_valid`` ``[B,pairs]``.  Returns (z_nodes, z_edges, ldj, reg_nodes [B], reg_edges [B])."""
    if ldj is None:
        ldj = z_nodes.new_zeros(z_nodes.size(0))
    mv = mask_valid.unsqueeze(-1)
    mn = mask_nodes[None, :min(mask_nodes.size(0), z_nodes.size(1)), :]                     # :52-53
    me = mask_edges[None, :min(mask_edges.size(0), z_edges.size(1)), :]
    res = []
    for z, nn_out, m, k, sf, msf, pd in ((z_nodes, nn_nodes, mn, k_nodes, sf_nodes, msf_nodes, pad),
                                         (z_edges, nn_edges, me, k_edges, sf_edges, msf_edges, mv)):
        p = mixt_params(nn_out * pd, m, k, sf, msf)                                           # :64,:78,:113-118
        out, l, reg = mixcdf_run(z, *p, reverse=reverse, mask=m, pad=pd, reg_max=reg_max, reg_factor=reg_factor,
                                 training=training)                                           # :119-135
        reg_b = reg.float().sum(dim=[1, 2]) if reg is not None else None                      # :138-139
        res.append((out.float() * pd, l.float(), reg_b))                                      # :74,:88,:136-137
    (zn, ln, rn), (ze, le, re) = res
    return zn, ze, ldj + ln + le, rn, re


def node_edge_wrapper(fn, z_nodes, z_edges, node_args, edge_args, *, ldj=None, reverse=False, length=None, pad=None,
                      mask_valid=None):
    """``NodeEdgeFlowWrapper.forward`` (graph_node_edge_coupling.py:158-165): ``fn`` = :func:`actnorm` or
    :func:`invconv` applied to the nodes with (length, pad) and to the edges with
    (edge_length = mask_valid.sum(1), mask_valid[..., None])."""
    zn, ldj = fn(z_nodes, *node_args, ldj=ldj, reverse=reverse, length=length, pad=pad)
    ze, ldj = fn(z_edges, *edge_args, ldj=ldj, reverse=reverse, length=mask_valid.sum(dim=1), pad=mask_valid.unsqueeze(-1))
    return zn, ze, ldj


# ---------------------------------------------------------------------------
# a5: affine coupling   (layers/flows/coupling_layer.py:42-98)
# ---------------------------------------------------------------------------
def affine_params(nn_out, mask, scaling_factor=None):
    """coupling_layer.py:76-85 - record per channel is ``[s, t]``."""
    rec = nn_out.view(nn_out.shape[:-1] + (nn_out.shape[-1] // 2, 2))
    s, t = rec[..., 0], rec[..., 1]
    if scaling_factor is not None:
        fac = scaling_factor.exp().view(1, 1, -1)
        s = torch.tanh(s / fac.clamp(min=1.0)) * fac
    return s * (1 - mask), t * (1 - mask)


def affine_coupling(z, nn_out, mask, scaling_factor, ldj=None, reverse=False):
    """coupling_layer.py:53-65, 88-98.  ldj is *not* pad-masked (App. B #4)."""
    if ldj is None:
        ldj = z.new_zeros(z.size(0))
    s, t = affine_params(nn_out, mask, scaling_factor)
    if not reverse:
        out, l = (z + t) * torch.exp(s), s.sum(dim=[1, 2])
    else:
        out, l = z * torch.exp(-s) - t, -s.sum(dim=[1, 2])
    return out, ldj + l


# ---------------------------------------------------------------------------
# a7 / a8: activation normalisation   (layers/flows/activation_normalization.py)
# ---------------------------------------------------------------------------
def actnorm(z, bias, scales, ldj=None, reverse=False, length=None, pad=None):
    """activation_normalization.py:24-48.  ``bias``/``scales`` are ``[1,1,C]``."""
    if ldj is None:
        ldj = z.new_zeros(z.size(0))
    if length is None:
        length = z.size(1) if pad is None else pad.squeeze(2).sum(dim=1)
    else:
        length = length.float()
    if not reverse:
        out = (z + bias) * torch.exp(scales)
        ldj = ldj + scales.sum(dim=[1, 2]) * length
    else:
        out = z * torch.exp(-scales) - bias
        ldj = ldj + (-scales.sum(dim=[1, 2])) * length
    if pad is not None:
        out = out * pad
    return out, ldj


def actnorm_data_init(x, pad=None):
    """activation_normalization.py:55-67 - returns (bias, scales) as ``[1,1,C]``."""
    m = pad if pad is not None else x.new_ones(x.shape)
    n = m.sum(dim=[0, 1], keepdim=True)
    bias = -(x * m).sum(dim=[0, 1], keepdim=True) / n
    var = (((x + bias) ** 2) * m).sum(dim=[0, 1], keepdim=True) / n
    return bias, -0.5 * var.log()


def ext_actnorm(z, bias, raw_scales, ldj=None, reverse=False, pad=None):
    """activation_normalization.py:116-144.  ``bias``/``raw_scales`` are the two
    halves of ``pred_net(ext_input)`` (``[B,S,D]`` each); the scale half is tanh-bounded."""
    if ldj is None:
        ldj = z.new_zeros(z.size(0))
    pm = 1.0 if pad is None else pad
    s = torch.tanh(raw_scales)
    if not reverse:
        out = (z + bias) * torch.exp(s)
        ldj = ldj + (s * pm).sum(dim=[1, 2])
    else:
        out = z * torch.exp(-s) - bias
        ldj = ldj - (s * pm).sum(dim=[1, 2])
    return out, ldj


# ---------------------------------------------------------------------------
# a9: invertible 1x1 convolution   (layers/flows/permutation_layers.py:61-136)
# ---------------------------------------------------------------------------
def invconv_weight(p, l, log_s, u, sign_s):
    """permutation_layers.py:67-71 - W = P (L o tril + I)(U o triu + diag(sign e^{log_s}))."""
    C = l.shape[0]
    tril = torch.tril(torch.ones(C, C, dtype=l.dtype), -1)
    lo = l * tril + torch.eye(C, dtype=l.dtype)
    up = u * tril.t() + torch.diag(sign_s * torch.exp(log_s))
    return p @ (lo @ up), log_s.sum()


def invconv_inverse(weight):
    """permutation_layers.py:77,85 - inverse taken in float64, rounded to float32."""
    return torch.inverse(weight.double()).float()


def invconv(z, weight, sldj, ldj=None, reverse=False, length=None, pad=None):
    """permutation_layers.py:106-121.  ``weight`` is W (forward) or W^-1 (reverse)."""
    if ldj is None:
        ldj = z.new_zeros(z.size(0))
    n = z.size(1) if length is None else length.float()
    ldj = ldj - sldj * n if reverse else ldj + sldj * n
    out = torch.matmul(z, weight.unsqueeze(0))
    if pad is not None:
        out = out * pad
    return out, ldj


# ---------------------------------------------------------------------------
# a11: logistic prior   (layers/flows/distributions.py:117-163)
# ---------------------------------------------------------------------------
def logistic_from_uniform(u, mu=0.0, sigma=LOGISTIC_SIGMA, eps=LOGISTIC_EPS):
    """distributions.py:139-145 + 117-127 - squeeze U(0,1) into (eps/2, 1-eps/2), logit in f64."""
    v = (u * (1 - eps)) + eps / 2
    v = v.double()
    return (-torch.log(v.reciprocal() - 1.0)).float() * sigma + mu


def logistic_log_prob(x, mu=0.0, sigma=LOGISTIC_SIGMA):
    """distributions.py:129-136, 154-163."""
    v = (x - mu) / sigma
    return -(F.softplus(v) + F.softplus(-v) + float(np.log(sigma)))


# ---------------------------------------------------------------------------
# a12: mixture-of-logistics categorical encoding (num_flows = 0)
#      (layers/categorical_encoding/linear_encoding.py:59-196)
# ---------------------------------------------------------------------------
def categ_table(embed_weight, lin_weight, lin_bias):
    """Per-class (bias, raw scale) rows: ``pred_net(embed(v))`` for every class v
    (linear_encoding.py:138; activation_normalization.py:127-128; help_layers.py:57-73).
    Returns ``[V, 2D]``."""
    return F.linear(embed_weight, lin_weight, lin_bias)


def categ_encode(x, u_noise, table, category_prior, beta=1.0, pad=None):
    """linear_encoding.py:71-92 + 153-174 for the mixture model (one ExtActNorm).

    x: ``[B,S]`` int64 tokens; u_noise: ``[B*S,1,D]`` U(0,1) draws that the
    reference would take from ``prior_distribution.sample``; table: ``[V,2D]``.
    Returns ``(z [B,S,D] f32, ldj [B] f32, class_prob_log [B*S])``.
    """
    B, S = x.shape
    V, D = table.shape[0], table.shape[1] // 2
    tok = x.reshape(B * S)
    padf = pad.reshape(B * S, 1, -1) if pad is not None else x.new_ones((B * S, 1, 1), dtype=torch.float32)
    z0 = logistic_from_uniform(u_noise)                                   # :76
    init_log_p = logistic_log_prob(z0).sum(dim=[1, 2])                    # :77
    b_tok = table[tok, :D].unsqueeze(1)
    s_tok = torch.tanh(table[tok, D:]).unsqueeze(1)
    z = (z0 + b_tok) * torch.exp(s_tok)                                   # ExtActNorm fwd
    ldj_fwd = s_tok.sum(dim=[1, 2])
    log_point = init_log_p - ldj_fwd + category_prior[tok]               # :82-83
    # reverse pass of every class-conditional flow (:155-160)
    b_all = table[:, :D].unsqueeze(0)
    s_all = torch.tanh(table[:, D:]).unsqueeze(0)
    z_back = z * torch.exp(-s_all) - b_all                                 # [B*S, V, D]
    back_log_p = logistic_log_prob(z_back).sum(dim=-1)
    denom = back_log_p - s_all.sum(dim=-1) + category_prior[None, :]      # :163-164
    own = F.one_hot(tok, V).to(denom.dtype)                               # :167-168
    denom = denom * (1 - own) + log_point.unsqueeze(-1) * own
    class_prob_log = log_point - torch.logsumexp(denom, dim=-1)           # :170-173
    ldj_loc = (beta * class_prob_log - (init_log_p - ldj_fwd)) * padf.squeeze()   # :89-90
    z = z * padf
    return z.reshape(B, S, D), ldj_loc.reshape(B, S).sum(dim=-1), class_prob_log


def categ_decode(z, table, category_prior):
    """linear_encoding.py:184-196 - argmax over classes of the class-conditional density."""
    B, S, D = z.shape
    zz = z.reshape(B * S, 1, D)
    b_all = table[:, :D].unsqueeze(0)
    s_all = torch.tanh(table[:, D:]).unsqueeze(0)
    z_back = zz * torch.exp(-s_all) - b_all
    score = logistic_log_prob(z_back).sum(dim=-1) - s_all.sum(dim=-1) + category_prior[None, :]
    return score.argmax(dim=-1).reshape(B, S)


# ---------------------------------------------------------------------------
# a13: linear decoder  (layers/categorical_encoding/decoder.py:35-63)
# ---------------------------------------------------------------------------
def decoder_features(z):
    """decoder.py:57 - [z, elu(z), elu(-z)] along the last axis."""
    return torch.cat([z, F.elu(z), F.elu(-z)], dim=-1)


# ---------------------------------------------------------------------------
# a10: the flow container's running ldj accumulator  (layers/flows/flow_model.py:25-53)
# ---------------------------------------------------------------------------
def bits_per_dim(ldj, log_prior, length):
    """general/task.py:148-149 with the loss of experiments/*/task.py: nats/element -> bits."""
    return ((-ldj - log_prior) / length).mean() * math.log2(math.e)


def lm_flow_forward(tokens, u_noise, enc, blocks, pad=None, length=None, beta=1.0):
    """Reference composition for BASELINE config 2 with the coupling nets' outputs given:
    encode -> n x [ActNorm, InvConv, MixtureCDFCoupling] -> prior log-prob.

    ``enc`` = dict(table, prior); each block = dict(bias, scales, weight, sldj, nn_fn | nn_out,
    mask, K, sf, msf).  ``nn_fn(z_in) -> nn_out`` stands in for the black-box coupling
    network (coupling_layer.py:28-39).  Returns (z, ldj [B], log_prior [B]).
    """
    z, ldj, _ = categ_encode(tokens, u_noise, enc["table"], enc["prior"], beta=beta, pad=pad)
    for blk in blocks:
        z, l = actnorm(z, blk["bias"], blk["scales"], ldj=None, length=length, pad=pad)
        ldj = ldj + l                                                       # flow_model.py:44
        z, l = invconv(z, blk["weight"], blk["sldj"], ldj=None, length=length, pad=pad)
        ldj = ldj + l
        m = expand_mask(blk["mask"], z)
        nn_out = blk["nn_fn"](z * m) if "nn_fn" in blk else blk["nn_out"]
        z, l, _ = mixcdf_coupling(z, nn_out, m, blk["K"], blk["sf"], blk["msf"], pad=pad, training=False)
        ldj = ldj + l
    lp = logistic_log_prob(z)
    if pad is not None:
        lp = lp * pad
    return z, ldj, lp.sum(dim=[1, 2])


# ---------------------------------------------------------------------------
# a12 with linear flows (BASELINE config 1: "4 affine couplings"): LinearCategoricalEncoding(num_flows > 0)
#   flows = num_flows x [ExtActNorm, InvertibleConv, affine CouplingLayer(LinearNet)]   (linear_encoding.py:224-256)
# Works on a state dict with the reference's parameter names.
# ---------------------------------------------------------------------------
def _sd_linear(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd[key + ".bias"])


def linear_net(sd, pre, x, ext_input=None):
    """LinearNet.forward (layers/networks/help_layers.py:76-102): inp_layer (Linear + GELU), external input concatenated,
    main_net = (Linear, GELU) x num_layers + Linear."""
    h = F.gelu(_sd_linear(sd, pre + "inp_layer.0", x))
    if ext_input is not None:
        h = torch.cat([h, ext_input], dim=-1)
    idx = sorted({int(k[len(pre + "main_net."):].split(".")[0]) for k in sd if k.startswith(pre + "main_net.")})
    for n, i in enumerate(idx):
        h = _sd_linear(sd, "%smain_net.%d" % (pre, i), h)
        if n < len(idx) - 1:
            h = F.gelu(h)
    return h


def encoding_flow_pass(sd, z, tokens, num_flows, reverse):
    """``_flow_forward`` (linear_encoding.py:134-141) on ``z`` [M,1,D] for class ids ``tokens`` [M,1] -> (z, ldj [M])."""
    D = z.shape[-1]
    embed = sd["embed_layer.weight"][tokens]                                   # [M,1,E]
    mask = expand_mask(channel_mask(D, 0.5), z)
    ldj = z.new_zeros(z.size(0))
    order = range(num_flows) if not reverse else reversed(range(num_flows))
    for f in order:
        a, c, p = 3 * f, 3 * f + 1, 3 * f + 2
        w, sldj = invconv_weight(*(sd["flow_layers.%d.%s" % (c, k)] for k in ("p", "l", "log_s", "u", "sign_s")))
        ext = _sd_linear(sd, "flow_layers.%d.pred_net.layer" % a, embed)
        bias, raw = ext.chunk(2, dim=2)

        def coupling(zz, ll):
            nn_out = linear_net(sd, "flow_layers.%d.nn." % p, zz * mask, ext_input=embed)
            out, l = affine_coupling(zz, nn_out, mask, sd["flow_layers.%d.scaling_factor" % p], reverse=reverse)
            return out, ll + l

        if not reverse:
            z, ldj = ext_actnorm(z, bias, raw, ldj)
            z, ldj = invconv(z, w, sldj, ldj)
            z, ldj = coupling(z, ldj)
        else:
            z, ldj = coupling(z, ldj)
            z, ldj = invconv(z, invconv_inverse(w), sldj, ldj, reverse=True)
            z, ldj = ext_actnorm(z, bias, raw, ldj, reverse=True)
    return z, ldj


def categ_encode_flows(sd, x, u_noise, num_flows, beta=1.0, pad=None):
    """LinearCategoricalEncoding.forward with linear flows (linear_encoding.py:59-132, 153-174).
    Returns (z [B,S,D], ldj [B])."""
    B, S = x.shape
    V = sd["category_prior"].shape[0]
    D = u_noise.shape[-1]
    tok = x.reshape(B * S, 1)
    padf = pad.reshape(B * S, 1, -1) if pad is not None else torch.ones(B * S, 1, 1)
    z0 = logistic_from_uniform(u_noise)                                         # prior of the encoding flows (:39)
    init_log_p = logistic_log_prob(z0).sum(dim=[1, 2])
    z, ldj_fwd = encoding_flow_pass(sd, z0, tok, num_flows, reverse=False)
    log_point = init_log_p - ldj_fwd + sd["category_prior"][tok.squeeze(-1)]
    z_all = z.expand(-1, V, -1).reshape(-1, 1, D)
    cls = torch.arange(V)[None, :].expand(B * S, -1).reshape(-1, 1)
    z_back, ldj_back = encoding_flow_pass(sd, z_all, cls, num_flows, reverse=True)
    back_log_p = logistic_log_prob(z_back).sum(dim=[1, 2])
    denom = (back_log_p + ldj_back).view(B * S, V) + sd["category_prior"][None, :]
    own = F.one_hot(tok.squeeze(-1), V).to(denom.dtype)
    denom = denom * (1 - own) + log_point.unsqueeze(-1) * own
    class_prob_log = log_point - torch.logsumexp(denom, dim=-1)
    ldj_loc = (beta * class_prob_log - (init_log_p - ldj_fwd)) * padf.squeeze()
    return (z * padf).reshape(B, S, D), ldj_loc.reshape(B, S).sum(dim=-1)


def categ_decode_flows(sd, z, num_flows):
    """``_posterior_sample`` (linear_encoding.py:184-196): arg-max class of the flow posterior."""
    B, S, D = z.shape
    V = sd["category_prior"].shape[0]
    z_all = z.reshape(B * S, 1, D).expand(-1, V, -1).reshape(-1, 1, D)
    cls = torch.arange(V)[None, :].expand(B * S, -1).reshape(-1, 1)
    z_back, ldj_back = encoding_flow_pass(sd, z_all, cls, num_flows, reverse=True)
    logp = (logistic_log_prob(z_back).sum(dim=[1, 2]) + ldj_back).view(B * S, V) + sd["category_prior"][None, :]
    return logp.argmax(dim=-1).reshape(B, S)


def decoder_linear(sd, z, pre=""):
    """DecoderLinear.forward (layers/categorical_encoding/decoder.py:56-60)."""
    return torch.log_softmax(linear_net(sd, pre + "layers.", decoder_features(z)), dim=-1)


# ---------------------------------------------------------------------------
# SigmoidFlow and VariationalDequantization
#   layers/flows/sigmoid_layer.py:24-48; layers/categorical_encoding/variational_dequantization.py:31-98
# ---------------------------------------------------------------------------
SIGMOID_ALPHA = 1e-5


def sigmoid_flow(z, ldj=None, reverse=False, reverse_layer=False, sum_ldj=True):
    """sigmoid_layer.py:24-48 (fp32, the reference's own operation order)."""
    if ldj is None:
        ldj = z.new_zeros(z.size(0))
    alpha = SIGMOID_ALPHA
    if reverse_layer == reverse:                                                # XOR false -> sigmoid direction (:29-33)
        layer_ldj = -z - 2 * F.softplus(-z)
        out = torch.sigmoid(z)
    else:
        y = z * (1 - alpha) + alpha * 0.5
        layer_ldj = -torch.log(y) - torch.log(1 - y) + math.log(1 - alpha)
        out = torch.log(y) - torch.log(1 - y)
    if sum_ldj:
        return out, ldj + layer_ldj.view(z.size(0), -1).sum(dim=1)
    return out, layer_ldj


def dequant_example_net(sd, pre, x, ext_input):
    """The coupling network of the reference's usage example (variational_dequantization.py:118-131): the user-supplied
    black box ``model_func`` - Linear(1,h) on the noise, concatenated with the embedding, Linear-ReLU-Linear."""
    h = torch.cat([_sd_linear(sd, pre + "inp_layer", x), ext_input], dim=-1)
    return _sd_linear(sd, pre + "main_net.2", F.relu(_sd_linear(sd, pre + "main_net.0", h)))


def variational_dequantization(sd, x, u_noise, num_flows, net=dequant_example_net):
    """VariationalDequantization.forward, reverse=False (variational_dequantization.py:31-52, 61-66, 76-98).
    ``sd``: the module's state dict; ``x`` [B,S] int64; ``u_noise`` [B,S] in [0,1).  Returns (z_out [B,S,1], ldj [B])."""
    r = u_noise.unsqueeze(-1)                                                   # fp32 upstream; fp64 for autograd checks
    ldj = torch.zeros(x.shape[0], dtype=r.dtype)
    r, ldj = sigmoid_flow(r, ldj, reverse=False, reverse_layer=True)
    embed = sd["embed_layer.weight"][x]
    for f in range(num_flows):
        pre_a, pre_c = "flow_layers.%d." % (2 * f), "flow_layers.%d." % (2 * f + 1)
        r, ldj = actnorm(r, sd[pre_a + "bias"], sd[pre_a + "scales"], ldj)
        mask = expand_mask(sd[pre_c + "mask"], r)
        nn_out = net(sd, pre_c + "nn.", r * mask, embed)
        r, ldj = affine_coupling(r, nn_out, mask, sd[pre_c + "scaling_factor"], ldj)
    r, ldj = sigmoid_flow(r, ldj, reverse=True, reverse_layer=True)
    return x.to(r.dtype).unsqueeze(-1) + r, ldj


def dequantization_reverse(z, vocab_size):
    """variational_dequantization.py:53-56."""
    return torch.floor(z).clamp(min=0, max=vocab_size - 1).long().squeeze(dim=-1)
